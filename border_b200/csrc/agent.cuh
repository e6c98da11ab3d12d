// agent.cuh -- the object behind the opaque bb_agent handle: Policy + Agent + SyncModel
// (border-core/src/base/{policy,agent}.rs, border-async-trainer/src/sync_model.rs).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "nn.cuh"
#include "replay_internal.cuh"

namespace bb {

// host-side fastrand (wyrand) used by the explorers, as in the reference (dqn/explorer.rs:71,82)
struct FastRand {
    uint64_t s;
    explicit FastRand(uint64_t seed = 0) : s(seed) {}
    uint64_t u64() {
        s += 0xA0761D6478BD642FULL;
        __uint128_t t = (__uint128_t)s * (__uint128_t)(s ^ 0xE7037ED1A0B428DBULL);
        return (uint64_t)t ^ (uint64_t)(t >> 64);
    }
    uint32_t u32() { return (uint32_t)u64(); }
    float f32() {
        uint32_t b = 0x3F800000u | (u32() >> 9);
        float f; memcpy(&f, &b, 4);
        return f - 1.0f;
    }
    double f64() {
        uint64_t b = 0x3FF0000000000000ULL | (u64() >> 12);
        double f; memcpy(&f, &b, 8);
        return f - 1.0;
    }
    uint32_t u32_below(uint32_t n) {
        uint32_t x = u32();
        uint64_t m = (uint64_t)x * n;
        uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
        if (lo < n) {
            uint32_t t = (0u - n) % n;
            while (lo < t) { x = u32(); m = (uint64_t)x * n; hi = (uint32_t)(m >> 32); lo = (uint32_t)m; }
        }
        return hi;
    }
    uint64_t u64_below(uint64_t n) {
        uint64_t x = u64();
        __uint128_t m = (__uint128_t)x * n;
        uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
        if (lo < n) {
            uint64_t t = (0ull - n) % n;
            while (lo < t) { x = u64(); m = (__uint128_t)x * n; hi = (uint64_t)(m >> 64); lo = (uint64_t)m; }
        }
        return hi;
    }
};

// One VarStore + optimizer (DqnModel dqn/model/base.rs:20-39, Critic, Actor, EntCoef).
struct Model {
    std::string name;
    std::vector<ParamInfo> params;
    size_t n = 0;  // floats (padded)
    float *p = nullptr, *g = nullptr, *m = nullptr, *v = nullptr;
    // p is followed by its lo plane (p_lo()[i] = p[i] - tf32_trunc(p[i]), the second operand plane of the TMA-fed GEMMs):
    // Adam and the Polyak update refresh it in the same kernel, host-side writers call refresh_lo()
    float* p_lo() const { return p + n; }
    long plane() const { return (long)n; }
    void refresh_lo(const Ctx& c);
    uint64_t step = 0;
    AdamHyper hyper{};
    bool has_opt = true;
    // floats behind g (same allocation, so the peers' CUDA-IPC mapping of g covers them): the receive area of the
    // flag-in-data gradient exchange (nn.cuh: grad_exchange_ll), 8 sender slots x 2 x g_ll_cap floats
    size_t g_ll_cap = 0;
    void alloc(bool with_opt);
    void release();
    void set_hyper(const bb_opt_cfg& o);
    const ParamInfo* find(const std::string& nm) const;
    void copy_params_from(const Model& src, cudaStream_t s);
};

struct Agent {
    int device = 0;
    Ctx ctx;
    Ctx side_ctx[2];  // side streams + workspaces for the concurrent branches of a step (Ctx::side)
    bool train = false;
    uint64_t n_opts = 0;
    std::vector<Model*> models;  // registered VarStores by name
    // pinned scratch for records / policy outputs
    float* h_scratch = nullptr;
    float* d_scratch = nullptr;
    // multi-GPU gradient exchange: peers' gradient buffers and barrier flags mapped
    // with CUDA IPC; the fused all-reduce + Adam kernel reads them over NVLink.
    int rank = 0, world = 1;
    const float* peer_grad[8] = {nullptr};
    unsigned int* peer_flag[8] = {nullptr};  // peer_flag[r] = rank r's flag array (8 slots)
    unsigned int* my_flags = nullptr;
    unsigned int sync_epoch = 0;
    const float* const* peer_grads() const { return world > 1 ? peer_grad : nullptr; }
    // overlapped exchange (nn.cuh: Exchange): own stream so that waiting for peers never blocks a compute stream
    Ctx comm_ctx;
    unsigned int* xchg_ctr = nullptr;
    Exchange exchange() const;
    // region 0 = [split, n) is exchanged from inside the backward pass (begin_early_exchange, on comm_ctx), region 1 =
    // [0, split) and the optimizer follow at the end (synced_adam with early = true)
    void begin_early_exchange(Model& m, size_t split);
    // [lo, hi) (gradients that exist before the end of the backward pass, but are small): flag-in-data exchange on comm_ctx
    void mid_exchange(Model& m, size_t lo, size_t hi);
    void join_early_exchange();
    void comm_follows_compute();
    void grad_sync_begin();  // all ranks' gradients complete before anyone reads them
    void grad_sync_end();    // everyone done reading before anyone overwrites
    // Optimizer step of a data-parallel replica.  world 1: plain Adam.  world 2-3: ONE kernel reads every rank's
    // gradient and applies Adam (fused all-reduce + optimizer).  world >= 4: sharded mean (reduce-scatter +
    // broadcast through peer stores) then local Adam.  BB_GRAD_SYNC=fused|sharded overrides the choice.
    void synced_adam(Model& m, bool early = false, size_t split = 0 /* start of the early region */, size_t mid = 0 /* [mid, split) went through mid_exchange */);

    virtual ~Agent();
    void init_base(int dev);
    Model* model(const std::string& name);
    virtual void opt(Replay& rb, bb_record* rec) = 0;
    virtual void sample(const void* obs, size_t n, void* act_out) = 0;
    // Sampler::sample_and_push (border-core/src/trainer/sampler.rs:99-144) with the observation crossing PCIe once:
    // see bb_actor_step in border_b200.h.  Discrete-action agents only.
    // The generic loop is Agent::actor_step_n; a discrete-action agent supplies its observation row size, action count, the
    // policy forward for device-resident observations and the explorer's host-side draws.
    static constexpr int kActorMaxEnvs = 8;                                       // environments per call (vectorised Sampler)
    struct ActorPick { int mode = 0; long long forced = 0; double u = 0.0; };   // 0 argmax, 1 forced action, 2 softmax on u
    virtual size_t actor_obs_row_bytes() const { return 0; }                     // 0 = no device-side actor path
    virtual int actor_n_actions() const { return 0; }
    // Q values [n][A] for n observations already in HBM (contiguous rows), on ctx.stream
    virtual const float* actor_q(const uint8_t* d_obs, int n) { (void)d_obs; (void)n; return nullptr; }
    // the explorer's host-side draws for n observations, in the order Policy::sample(obs[n]) makes them
    virtual void actor_pick(int n, ActorPick* out) { for (int i = 0; i < n; ++i) out[i] = ActorPick{}; }
    // n environments at once: obs [n][row]; reset_obs [n][row] + reset_mask[n] (or both null); reward / term / trunc [n]
    void actor_step_n(Replay& rb, int n, const void* obs, const void* reset_obs, const int8_t* reset_mask, const float* reward,
                      const int8_t* term, const int8_t* trunc, int64_t* act_out, bool obs_on_device);
    void actor_step(Replay& rb, const void* obs, const void* reset_obs, float reward, int8_t term, int8_t trunc,
                    int64_t* act_out, bool obs_on_device = false) {
        const int8_t one = 1;
        actor_step_n(rb, 1, obs, reset_obs, reset_obs ? &one : nullptr, &reward, &term, &trunc, act_out, obs_on_device);
    }
    void actor_reset() { actor_has_prev = false; }
    // three rotating observation buffers ([kActorMaxEnvs][row] each): the previous step's acting observations (obs of the
    // transitions being pushed), this step's uploaded observations (their next_obs), and -- only when an episode ended --
    // the acting observations of this step (uploaded rows with the reset observations written over them)
    uint8_t* d_actor_obs[3] = {nullptr, nullptr, nullptr};
    uint8_t* d_actor_misc = nullptr;    // reward f32[8] | term i8[8] | trunc i8[8]
    uint8_t* h_actor_stage = nullptr;   // pinned: obs rows | reset rows | misc
    long long* d_actor_act = nullptr;   // [kActorMaxEnvs]
    long long* h_actor_act = nullptr;   // pinned, written by the explorer kernel
    size_t actor_row_pad = 0;
    int actor_prev = 0, actor_n = 0;
    bool actor_has_prev = false;
    cudaEvent_t ev_actor = nullptr;
    virtual Model* sync_model_src() = 0;  // which VarStore SyncModel ships (DQN qnet, SAC pi)
    virtual void inject_noise(int slot, const float* host, size_t n);
    virtual void precision_changed() {}   // drop captured graphs (they hold the GEMM kernels of the previous precision mode)
    virtual void grad_buffer(void** p, uint64_t* n);
    void save_params(const char* dir);
    void load_params(const char* dir);
};

// One update step as a CUDA graph (what Dqn::update_critic does by hand, for the other agents): after three eager updates
// the launches of `enqueue` are captured once per (replay, batch, stream) and replayed.  Everything `enqueue` launches must
// be argument-invariant: step-dependent scalars live in device memory (AdamScalars), the replay draws from its device-side
// position.  `advance` does the host-side bookkeeping of a replayed sample (rb.sample(B, &bv, false)).
struct UpdateGraph {
    cudaGraphExec_t exec = nullptr;
    const void* key[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t eager = 0, kernels = 0;
    bool broken = false;
    void reset() { if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; } }
    ~UpdateGraph() { reset(); }
    template <class Enqueue, class Advance>
    void run(const Ctx& ctx, Replay& rb, int B, bool allowed, Enqueue enqueue, Advance advance) {
        const char* genv = getenv("BB_GRAPH");  // read per call so tests can flip it
        const bool want = allowed && !(genv && atoi(genv) == 0) && !broken && !ctx.prof && !rb.per && ctx.stream != nullptr &&
                          ctx.stream != cudaStreamLegacy && ctx.stream != cudaStreamPerThread && rb.stream == ctx.stream &&
                          eager >= 3 && rb.batch_cap >= (size_t)B;
        if (want) {
            const void* k[4] = {&rb, (const void*)(uintptr_t)B, (const void*)ctx.stream,
                                (const void*)((uintptr_t)rb.b_obs ^ (uintptr_t)(rb.batch_generation << 48))};
            if (exec && memcmp(k, key, sizeof(k)) == 0) {
                advance();
                BB_CUDA(cudaGraphLaunch(exec, ctx.stream));
                g_launch_count.fetch_add(kernels, std::memory_order_relaxed);
                return;
            }
            reset();
            const uint64_t rng0 = rb.rng_pos;
            const size_t lb0 = rb.last_batch;
            const uint64_t n0 = g_launch_count.load();
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                try { enqueue(); } catch (...) { ok = false; }
                if (cudaStreamEndCapture(ctx.stream, &graph) != cudaSuccess || !graph) ok = false;
            }
            if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { ok = false; exec = nullptr; }
            if (graph) cudaGraphDestroy(graph);
            if (ok) {
                kernels = g_launch_count.load() - n0;
                memcpy(key, k, sizeof(k));
                BB_CUDA(cudaGraphLaunch(exec, ctx.stream));  // the capture enqueued nothing: run this update now
                return;
            }
            cudaGetLastError();  // clear the sticky capture error, fall back to eager launches for good
            broken = true;
            rb.rng_pos = rng0; rb.last_batch = lb0;
            g_launch_count.store(n0);
        }
        enqueue();
        eager += 1;
    }
};

Agent* make_dqn(const bb_dqn_cfg& cfg);
Agent* make_sac(const bb_sac_cfg& cfg);
Agent* make_iqn(const bb_iqn_cfg& cfg);

}  // namespace bb
