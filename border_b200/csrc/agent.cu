// agent.cu -- Agent/Model plumbing and the agent half of the C ABI.
#include <errno.h>
#include <string.h>
#include <sys/stat.h>
#include <fstream>
#include <mutex>
#include <sstream>
#include <stdlib.h>
#include "agent.cuh"

namespace bb {

// ------------------------------------------------------------------------------- Model

void Model::alloc(bool with_opt) {
    has_opt = with_opt;
    p = dev_alloc_zero<float>(2 * n);  // parameters + their lo plane
    if (with_opt) {
        g = dev_alloc_zero<float>(n + 16 * g_ll_cap);
        m = dev_alloc_zero<float>(n);
        v = dev_alloc_zero<float>(n);
    }
}
void Model::release() {
    cudaFree(p); cudaFree(g); cudaFree(m); cudaFree(v);
    p = g = m = v = nullptr;
}
void Model::set_hyper(const bb_opt_cfg& o) {
    if (o.kind == BB_OPT_ADAMW) {  // opt.rs:38-54
        hyper = AdamHyper{o.lr, o.beta1, o.beta2, o.eps, o.wd, true};
        BB_CHECK(!o.amsgrad, "AdamW amsgrad=true is not supported");
    } else {  // Adam::default(), opt.rs:33-36
        hyper = AdamHyper{o.lr, 0.9, 0.999, 1e-8, 0.0, false};
    }
}
const ParamInfo* Model::find(const std::string& nm) const {
    for (auto& pi : params)
        if (pi.name == nm) return &pi;
    return nullptr;
}
void Model::copy_params_from(const Model& src, cudaStream_t s) {
    BB_CHECK(src.n == n, "VarStore size mismatch");
    BB_CUDA(cudaMemcpyAsync(p, src.p, 2 * n * sizeof(float), cudaMemcpyDeviceToDevice, s));  // both planes
}
void Model::refresh_lo(const Ctx& c) { make_lo(c, p, p_lo(), n); }

// ------------------------------------------------------------------------------- Agent base

void Agent::init_base(int dev) {
    device = dev;
    int count = 0;
    BB_CUDA(cudaGetDeviceCount(&count));
    BB_CHECK(dev >= 0 && dev < count, "no such CUDA device");
    DeviceGuard g(dev);
    ctx.device = dev;
    ctx.sms = num_sms(dev);
    ctx.stream = device_stream(dev);
    ctx.alloc_scratch(8u << 20);  // 32 MB split-K / reduction workspace
    BB_CUDA(cudaMallocHost(&h_scratch, 4096 * sizeof(float)));
    d_scratch = dev_alloc<float>(1 << 20);
    BB_CUDA(cudaEventCreateWithFlags(&ctx.ev, cudaEventDisableTiming));
    static const bool serial = getenv("BB_SERIAL") && atoi(getenv("BB_SERIAL")) != 0;
    for (int i = 0; i < 2; ++i) {
        Ctx& s = side_ctx[i];
        s.device = dev; s.sms = ctx.sms;
        BB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        {
            cudaStream_t own = s.stream;
            s.stream = ctx.stream;  // zero on the main stream (synchronised below)
            s.alloc_scratch(ctx.ws_floats);
            s.stream = own;
        }
        BB_CUDA(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
        if (!serial) ctx.side[i] = &s;
    }
    comm_ctx.device = dev; comm_ctx.sms = ctx.sms;
    BB_CUDA(cudaStreamCreateWithFlags(&comm_ctx.stream, cudaStreamNonBlocking));
    BB_CUDA(cudaEventCreateWithFlags(&comm_ctx.ev, cudaEventDisableTiming));
    xchg_ctr = dev_alloc_zero<unsigned int>(32, ctx.stream);
    BB_CUDA(cudaStreamSynchronize(ctx.stream));
}
Agent::~Agent() {
    if (comm_ctx.stream) { cudaStreamSynchronize(comm_ctx.stream); cudaStreamDestroy(comm_ctx.stream); }
    if (comm_ctx.ev) cudaEventDestroy(comm_ctx.ev);
    cudaFree(xchg_ctr);
    cudaFree(d_actor_obs[0]); cudaFree(d_actor_obs[1]); cudaFree(d_actor_obs[2]); cudaFree(d_actor_act); cudaFree(d_actor_misc);
    if (h_actor_stage) cudaFreeHost(h_actor_stage);
    if (h_actor_act) cudaFreeHost(h_actor_act);
    if (ev_actor) cudaEventDestroy(ev_actor);
    for (int i = 0; i < 2; ++i) {
        if (side_ctx[i].stream) { cudaStreamSynchronize(side_ctx[i].stream); cudaStreamDestroy(side_ctx[i].stream); }
        side_ctx[i].free_scratch();
        if (side_ctx[i].ev) cudaEventDestroy(side_ctx[i].ev);
    }
    if (ctx.ev) cudaEventDestroy(ctx.ev);
    ctx.free_scratch();
    cudaFree(d_scratch);
    cudaFree(my_flags);
    if (h_scratch) cudaFreeHost(h_scratch);
}
Model* Agent::model(const std::string& name) {
    for (auto* m : models)
        if (m->name == name) return m;
    throw Error("no such model (VarStore): " + name);
}
void Agent::inject_noise(int, const float*, size_t) { throw Error("this agent takes no injected noise"); }
// Explorer in the tail of the policy forward (dqn/explorer.rs:29-31,68-90, iqn/explorer.rs:78-97; eval: dqn/base.rs:229-236).  The fastrand draws
// are made on the host in the reference's order and arrive as (mode, forced, u): 0 = argmax Q, 1 = the random action
// `forced`, 2 = softmax(Q).multinomial(1) by inverse CDF on u.  The action goes to device memory (the next push reads it)
// and to pinned host memory (the env reads it).
struct SelectParams {
    const float* q; int A; int n;
    int mode[Agent::kActorMaxEnvs]; long long forced[Agent::kActorMaxEnvs]; double u[Agent::kActorMaxEnvs];
    long long* act_dev; long long* act_host;
};
__global__ void actor_select_kernel(SelectParams s) {
    const int i = threadIdx.x;
    if (i >= s.n) return;
    const float* q = s.q + (size_t)i * s.A;
    long long a = 0;
    if (s.mode[i] == 1) {
        a = s.forced[i];
    } else if (s.mode[i] == 0) {
        int best = 0;
        for (int j = 1; j < s.A; ++j)
            if (q[j] > q[best]) best = j;
        a = best;
    } else {
        float mx = q[0];
        for (int j = 1; j < s.A; ++j) mx = fmaxf(mx, q[j]);
        double z = 0;
        for (int j = 0; j < s.A; ++j) z += exp((double)(q[j] - mx));
        const double u = s.u[i] * z;
        double acc = 0;
        int pick = s.A - 1;
        for (int j = 0; j < s.A; ++j) {
            acc += exp((double)(q[j] - mx));
            if (u < acc) { pick = j; break; }
        }
        a = pick;
    }
    s.act_dev[i] = a;
    reinterpret_cast<volatile long long*>(s.act_host)[i] = a;
}

void Agent::grad_buffer(void** p, uint64_t* n) { *p = nullptr; *n = 0; }

// bb_actor_step[_n] (border_b200.h): Sampler::sample_and_push with device-resident observations, n environments per call
void Agent::actor_step_n(Replay& rb, int n, const void* obs, const void* reset_obs, const int8_t* reset_mask, const float* reward,
                         const int8_t* term, const int8_t* trunc, int64_t* act_out, bool obs_on_device) {
    DeviceGuard g(device);
    const size_t row = actor_obs_row_bytes();
    BB_CHECK(row > 0, "bb_actor_step: this agent has no device-side actor path (discrete-action agents only: DQN, IQN)");
    BB_CHECK(n >= 1 && n <= kActorMaxEnvs, "bb_actor_step_n: 1..8 environments per call");
    BB_CHECK(rb.obs_row_bytes == row && rb.cfg.act_kind == BB_I64 && rb.cfg.act_elems == 1,
             "bb_actor_step: the replay rows do not match the network (obs row) / a scalar i64 action");
    BB_CHECK(!actor_has_prev || n == actor_n, "bb_actor_step_n: the number of environments changed (call bb_actor_reset first)");
    BB_CHECK((reset_obs == nullptr) == (reset_mask == nullptr), "bb_actor_step_n: reset_obs and reset_mask go together");
    const size_t block = (size_t)kActorMaxEnvs * row;   // one observation buffer
    if (!d_actor_act) {
        actor_row_pad = (block + 255) / 256 * 256;
        for (int k = 0; k < 3; ++k) d_actor_obs[k] = dev_alloc<uint8_t>(actor_row_pad);
        d_actor_misc = dev_alloc_zero<uint8_t>(64);
        d_actor_act = dev_alloc_zero<long long>(kActorMaxEnvs);
        BB_CUDA(cudaMallocHost(&h_actor_stage, 2 * actor_row_pad + 64));
        BB_CUDA(cudaMallocHost(&h_actor_act, kActorMaxEnvs * sizeof(long long)));
        BB_CUDA(cudaEventCreateWithFlags(&ev_actor, cudaEventDisableTiming));
    }
    const int prev = actor_prev, cur = (actor_prev + 1) % 3, alt = (actor_prev + 2) % 3;
    // this step's observations (the transitions' next_obs) -> `cur`; reward / flags -> the misc block
    uint8_t* h = h_actor_stage;
    uint8_t* hm = h_actor_stage + 2 * actor_row_pad;
    memcpy(hm, reward, 4 * (size_t)n);
    memcpy(hm + 32, term, (size_t)n);
    memcpy(hm + 40, trunc, (size_t)n);
    if (obs_on_device) {   // e.g. the frame stack bb_atari_step left in HBM: only reward + flags cross PCIe
        BB_CUDA(cudaMemcpyAsync(d_actor_obs[cur], obs, (size_t)n * row, cudaMemcpyDeviceToDevice, ctx.stream));
    } else {
        memcpy(h, obs, (size_t)n * row);
        BB_CUDA(cudaMemcpyAsync(d_actor_obs[cur], h, (size_t)n * row, cudaMemcpyHostToDevice, ctx.stream));
    }
    if (actor_has_prev) {
        BB_CUDA(cudaMemcpyAsync(d_actor_misc, hm, 48, cudaMemcpyHostToDevice, ctx.stream));
        if (rb.stream != ctx.stream) stream_wait(rb.stream, ctx.stream);
        rb.push(d_actor_obs[prev], d_actor_act, d_actor_obs[cur], (const float*)d_actor_misc, (const int8_t*)(d_actor_misc + 32),
                (const int8_t*)(d_actor_misc + 40), (size_t)n, true);
        if (rb.stream != ctx.stream) stream_wait(ctx.stream, rb.stream);
    }
    // the observations the policy acts on: the uploaded ones, with the reset observation in place of every finished
    // episode's last one (sampler.rs:128-137) -- a copy only when some episode ended
    int act_src = cur;
    bool any_reset = false;
    for (int i = 0; reset_mask && i < n; ++i) any_reset |= reset_mask[i] != 0;
    if (any_reset) {
        BB_CUDA(cudaMemcpyAsync(d_actor_obs[alt], d_actor_obs[cur], (size_t)n * row, cudaMemcpyDeviceToDevice, ctx.stream));
        for (int i = 0; i < n; ++i) {
            if (!reset_mask[i]) continue;
            const uint8_t* src = (const uint8_t*)reset_obs + (size_t)i * row;
            if (!obs_on_device) { memcpy(h + actor_row_pad + (size_t)i * row, src, row); src = h + actor_row_pad + (size_t)i * row; }
            BB_CUDA(cudaMemcpyAsync(d_actor_obs[alt] + (size_t)i * row, src, row,
                                    obs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx.stream));
        }
        act_src = alt;
    }
    const float* q = actor_q(d_actor_obs[act_src], n);
    ActorPick picks[kActorMaxEnvs];
    actor_pick(n, picks);
    SelectParams sp{};
    sp.q = q; sp.A = actor_n_actions(); sp.n = n; sp.act_dev = d_actor_act; sp.act_host = h_actor_act;
    for (int i = 0; i < n; ++i) { sp.mode[i] = picks[i].mode; sp.forced[i] = picks[i].forced; sp.u[i] = picks[i].u; }
    actor_select_kernel<<<1, 32, 0, ctx.stream>>>(sp);
    BB_LAUNCHED();
    ctx.phase = "policy"; ctx.layer = "explorer"; ctx.mark("actor_select");
    BB_CUDA(cudaEventRecord(ev_actor, ctx.stream));
    BB_CUDA(cudaEventSynchronize(ev_actor));
    for (int i = 0; i < n; ++i) act_out[i] = (int64_t)h_actor_act[i];
    actor_prev = act_src;
    actor_n = n;
    actor_has_prev = true;
}

// Cross-GPU barrier on flags in peer memory: rank r writes its epoch into slot r of every peer's
// flag array (system-scope store over NVLink), then waits until all slots of its own array reach
// the epoch.  Bounded spin: a missing peer cannot hang the GPU.
__global__ void peer_barrier_kernel(unsigned int* mine, unsigned int* const* peers_unused, unsigned int* p0,
                                    unsigned int* p1, unsigned int* p2, unsigned int* p3, unsigned int* p4,
                                    unsigned int* p5, unsigned int* p6, unsigned int* p7, int rank, int world,
                                    unsigned int epoch, int* err, long long timeout_cycles) {
    (void)peers_unused;
    unsigned int* peers[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
    int t = threadIdx.x;
    if (t < world) {
        __threadfence_system();
        volatile unsigned int* dst = peers[t] + rank;
        *dst = epoch;
        __threadfence_system();
        volatile unsigned int* src = mine + t;
        long long t0 = clock64();
        while ((int)(*src - epoch) < 0) {
            if (clock64() - t0 > timeout_cycles) {
                // a peer never arrived: the gradients this rank is about to read are not complete.  Raise the sticky
                // failure flag (every following opt() throws) instead of falling through to the optimizer silently.
                *reinterpret_cast<volatile int*>(err) = 21;
                __threadfence_system();
                break;
            }
        }
    }
}

// BB_PEER_TIMEOUT_S (default 30): how long a rank may lag (checkpointing, evaluation, graph capture on a peer) before the
// barrier gives up and raises the failure flag
static long long peer_timeout_cycles() {
    static const long long c = (long long)(getenv("BB_PEER_TIMEOUT_S") ? atof(getenv("BB_PEER_TIMEOUT_S")) : 30.0) * 1900000000LL;
    return c;
}

static void peer_barrier(Agent* a) {
    a->sync_epoch += 1;
    peer_barrier_kernel<<<1, 32, 0, a->ctx.stream>>>(a->my_flags, nullptr, a->peer_flag[0], a->peer_flag[1],
                                                    a->peer_flag[2], a->peer_flag[3], a->peer_flag[4], a->peer_flag[5],
                                                    a->peer_flag[6], a->peer_flag[7], a->rank, a->world, a->sync_epoch,
                                                    device_error_flag(), peer_timeout_cycles());
    BB_LAUNCHED();
}
void Agent::grad_sync_begin() {
    if (world > 1) peer_barrier(this);
}
void Agent::grad_sync_end() {
    if (world > 1) peer_barrier(this);
}
Exchange Agent::exchange() const {
    Exchange x{};
    for (int r = 0; r < 8; ++r) { x.grads[r] = r < world ? peer_grad[r] : nullptr; x.flags[r] = r < world ? peer_flag[r] : nullptr; }
    x.ctr = xchg_ctr; x.rank = rank; x.world = world; x.err = device_error_flag(); x.timeout_cycles = peer_timeout_cycles();
    return x;   // (ll_off / ll_cap: set by synced_adam from the model)
}

// Called from inside Net::backward once the weight gradients of the region's layers are enqueued (main stream + both side
// streams): the comm stream waits for all three and runs the region-0 exchange while the convolution backward continues.
void Agent::comm_follows_compute() {
    for (const Ctx* s : {(const Ctx*)&ctx, ctx.side[0], ctx.side[1]}) {
        if (!s) continue;
        BB_CUDA(cudaEventRecord(s->ev, s->stream));
        BB_CUDA(cudaStreamWaitEvent(comm_ctx.stream, s->ev, 0));
    }
    comm_ctx.prof = ctx.prof; comm_ctx.phase = "optimizer";
}
static bool ll_enabled() { const char* ll = getenv("BB_XCHG_LL"); return !(ll && !strcmp(ll, "0")); }
void Agent::mid_exchange(Model& m, size_t lo, size_t hi) {
    if (world <= 1 || hi <= lo) return;
    comm_follows_compute();
    Exchange x = exchange();
    x.ll_off = m.n; x.ll_cap = m.g_ll_cap;
    grad_exchange_ll(comm_ctx, x, lo, hi, 1);
}
void Agent::begin_early_exchange(Model& m, size_t split) {
    if (world <= 1) return;
    comm_follows_compute();
    static const int early_blocks = getenv("BB_XCHG_EARLY_BLOCKS") ? atoi(getenv("BB_XCHG_EARLY_BLOCKS")) : 32;
    grad_exchange(comm_ctx, exchange(), split, m.n, 0, early_blocks);
}
void Agent::join_early_exchange() {
    if (world <= 1) return;
    BB_CUDA(cudaEventRecord(comm_ctx.ev, comm_ctx.stream));
    BB_CUDA(cudaStreamWaitEvent(ctx.stream, comm_ctx.ev, 0));
}

// Optimizer step of a data-parallel replica.  world 1: plain Adam.  Otherwise the mean gradient is formed by
// grad_exchange_kernel (each rank reduces 1/world of a region and stores it to every rank over NVLink; the rendezvous is
// folded into the kernels) and Adam waits in its prologue for every rank's slice: with `early` the big region [split, n) was
// already exchanged under the backward pass (begin_early_exchange) and only [0, split) is exchanged here.
// BB_GRAD_SYNC=legacy keeps the round-1 sequence (barrier launch, reduce-scatter or all-read Adam, barrier launch).
void Agent::synced_adam(Model& m, bool early, size_t split, size_t mid) {
    if (world <= 1) {
        adam_step(ctx, m.p, m.g, m.m, m.v, m.n, m.hyper, m.step, nullptr, 1, m.p_lo());
        return;
    }
    const char* e = getenv("BB_GRAD_SYNC");
    if (e && (!strcmp(e, "legacy") || !strcmp(e, "sharded") || !strcmp(e, "fused"))) {
        BB_CHECK(!early, "the legacy gradient exchange has no early region");
        const bool sharded = !strcmp(e, "legacy") ? world >= 4 : !strcmp(e, "sharded");
        peer_barrier(this);  // every rank's gradient is complete
        if (sharded) {
            grad_reduce_scatter(ctx, peer_grad, m.n, rank, world);
            peer_barrier(this);  // every slice of the mean has landed in this rank's buffer (and nobody still reads the old one)
            adam_step(ctx, m.p, m.g, m.m, m.v, m.n, m.hyper, m.step, nullptr, 1, m.p_lo());
        } else {
            adam_step(ctx, m.p, m.g, m.m, m.v, m.n, m.hyper, m.step, peer_grad, world, m.p_lo());
            peer_barrier(this);  // everyone done reading before anyone overwrites
        }
        return;
    }
    Exchange x = exchange();
    x.ll_off = m.n; x.ll_cap = m.g_ll_cap;
    int regions = 1;
    if (early) {
        const size_t hi = mid > 0 ? mid : split;   // what is left for the end of the step
        if (hi > 0 && split <= m.g_ll_cap && ll_enabled()) grad_exchange_ll(ctx, x, 0, hi, 0);   // complete when the kernel ends
        else if (hi > 0) { BB_CHECK(mid == 0, "mid_exchange needs the flag-in-data path"); grad_exchange(ctx, x, 0, hi, 1); regions = 3; }
    } else {
        grad_exchange(ctx, x, 0, m.n, 0);
    }
    adam_step(ctx, m.p, m.g, m.m, m.v, m.n, m.hyper, m.step, nullptr, 1, m.p_lo(), &x, regions);
}

// ------------------------------------------------------------------------------- checkpoints

// One directory per save_params call, as in the reference (dqn/base.rs:348-362 writes
// qnet.pt.tch / qnet_tgt.pt.tch).  Each VarStore becomes `<model>.pt.tch.b200`: a text manifest
// line per tensor followed by raw little-endian f32 in the REFERENCE layout, plus (what the
// reference lacks) the Adam moments and step for true resume.
void Agent::save_params(const char* dir) {
    DeviceGuard g(device);
    BB_CUDA(cudaStreamSynchronize(ctx.stream));
    check_device_error("save_params");   // never checkpoint a model a timed-out kernel may have corrupted
    {   // fs::create_dir_all(&path) (dqn/base.rs:350): every missing component
        std::string d(dir);
        for (size_t i = 1; i <= d.size(); ++i)
            if (i == d.size() || d[i] == '/') {
                std::string part = d.substr(0, i);
                if (mkdir(part.c_str(), 0777) != 0 && errno != EEXIST) throw Error("cannot create directory " + part);
            }
    }
    for (Model* m : models) {
        std::string path = std::string(dir) + "/" + m->name + ".pt.tch.b200";
        std::ofstream f(path, std::ios::binary);
        if (!f) throw Error("cannot open " + path);
        std::vector<float> hp(m->n), hm, hv;
        BB_CUDA(cudaMemcpy(hp.data(), m->p, m->n * 4, cudaMemcpyDeviceToHost));
        if (m->has_opt) {
            hm.resize(m->n); hv.resize(m->n);
            BB_CUDA(cudaMemcpy(hm.data(), m->m, m->n * 4, cudaMemcpyDeviceToHost));
            BB_CUDA(cudaMemcpy(hv.data(), m->v, m->n * 4, cudaMemcpyDeviceToHost));
        }
        f << "BORDER_B200_VARSTORE 1 " << m->params.size() << " " << (m->has_opt ? 1 : 0) << " " << m->step << "\n";
        for (auto& pi : m->params) {
            f << pi.name << " " << pi.shape.size();
            for (auto d : pi.shape) f << " " << d;
            f << "\n";
        }
        std::vector<float> ref;
        for (auto& pi : m->params) {
            ref.resize(pi.numel);
            param_to_reference(pi, hp.data() + pi.offset, ref.data());
            f.write((const char*)ref.data(), pi.numel * 4);
            if (m->has_opt) {
                param_to_reference(pi, hm.data() + pi.offset, ref.data());
                f.write((const char*)ref.data(), pi.numel * 4);
                param_to_reference(pi, hv.data() + pi.offset, ref.data());
                f.write((const char*)ref.data(), pi.numel * 4);
            }
        }
        if (!f) throw Error("write failed: " + path);
    }
}

void Agent::load_params(const char* dir) {
    DeviceGuard g(device);
    BB_CUDA(cudaStreamSynchronize(ctx.stream));
    for (Model* m : models) {
        std::string path = std::string(dir) + "/" + m->name + ".pt.tch.b200";
        std::ifstream f(path, std::ios::binary);
        if (!f) throw Error("cannot open " + path);
        std::string magic;
        int ver, has_opt;
        size_t nt;
        uint64_t step;
        f >> magic >> ver >> nt >> has_opt >> step;
        if (magic != "BORDER_B200_VARSTORE" || nt != m->params.size()) throw Error("bad checkpoint " + path);
        if (ver != 1) throw Error("unsupported checkpoint version in " + path);
        for (auto& pi : m->params) {
            std::string nm;
            size_t nd;
            f >> nm >> nd;
            if (nm != pi.name || nd != pi.shape.size()) throw Error("checkpoint tensor mismatch: " + nm);
            for (size_t i = 0; i < nd; ++i) {
                int64_t d;
                f >> d;
                if (d != pi.shape[i]) throw Error("checkpoint shape mismatch: " + nm);
            }
        }
        f.get();  // newline
        std::vector<float> hp(m->n, 0.f), hm(m->n, 0.f), hv(m->n, 0.f), ref;
        for (auto& pi : m->params) {
            ref.resize(pi.numel);
            f.read((char*)ref.data(), pi.numel * 4);
            param_to_internal(pi, ref.data(), hp.data() + pi.offset);
            if (has_opt) {
                f.read((char*)ref.data(), pi.numel * 4);
                param_to_internal(pi, ref.data(), hm.data() + pi.offset);
                f.read((char*)ref.data(), pi.numel * 4);
                param_to_internal(pi, ref.data(), hv.data() + pi.offset);
            }
        }
        if (!f) throw Error("read failed: " + path);
        h2d_sync(m->p, hp.data(), m->n * 4, ctx.stream);
        m->refresh_lo(ctx);
        if (m->has_opt && has_opt) {
            h2d_sync(m->m, hm.data(), m->n * 4, ctx.stream);
            h2d_sync(m->v, hv.data(), m->n * 4, ctx.stream);
            m->step = step;
        } else if (m->has_opt) {
            // the file carries no optimizer state: start Adam afresh rather than keep the live model's stale moments
            BB_CUDA(cudaMemsetAsync(m->m, 0, m->n * 4, ctx.stream));
            BB_CUDA(cudaMemsetAsync(m->v, 0, m->n * 4, ctx.stream));
            BB_CUDA(cudaStreamSynchronize(ctx.stream));
            m->step = 0;
        }
    }
}

}  // namespace bb

// =============================================================================== C ABI

struct bb_agent { std::unique_ptr<bb::Agent> impl; };

static bb::Agent& A(bb_agent* a) {
    if (!a || !a->impl) throw bb::Error("null agent handle");
    return *a->impl;
}

extern "C" {

static void net_default(bb_net_cfg* n) {
    memset(n, 0, sizeof(*n));
    n->kind = BB_NET_MLP; n->in_dim = 4; n->n_units = 2; n->units[0] = 64; n->units[1] = 64; n->out_dim = 2;
    n->n_stack = 4;
}
static void opt_default(bb_opt_cfg* o, double lr) {
    memset(o, 0, sizeof(*o));
    o->kind = BB_OPT_ADAM; o->lr = lr; o->beta1 = 0.9; o->beta2 = 0.999; o->eps = 1e-8; o->wd = 0.0;
}

void bb_dqn_cfg_default(bb_dqn_cfg* c) {  // dqn/config.rs:82-102
    memset(c, 0, sizeof(*c));
    net_default(&c->q_config);
    opt_default(&c->opt_config, 0.0);
    c->soft_update_interval = 1; c->n_updates_per_opt = 1; c->batch_size = 1; c->discount_factor = 0.99;
    c->tau = 0.005; c->train = 0; c->explorer = BB_EXPLORER_SOFTMAX; c->eps_start = 1.0; c->eps_final = 0.02;
    c->final_step = 100000; c->double_dqn = 0; c->critic_loss = BB_LOSS_MSE; c->record_verbose_level = 0;
    c->device = 0; c->init_seed = 0; c->explorer_seed = 0x0123456789abcdefULL;
}
void bb_sac_cfg_default(bb_sac_cfg* c) {  // sac/config.rs:85-105
    memset(c, 0, sizeof(*c));
    net_default(&c->pi_config);
    net_default(&c->q_config);
    opt_default(&c->pi_opt_config, 3e-4);
    opt_default(&c->q_opt_config, 3e-4);
    c->gamma = 0.99; c->tau = 0.005; c->ent_coef_mode = BB_ENTCOEF_FIX; c->ent_coef_fix = 1.0;
    c->epsilon = 1e-4; c->min_lstd = -20.0; c->max_lstd = 2.0; c->n_updates_per_opt = 1; c->batch_size = 1;
    c->train = 0; c->critic_loss = BB_LOSS_MSE; c->reward_scale = 1.0; c->n_critics = 1; c->device = 0;
    c->noise_seed = 0x5ac5ac5acULL;
}
void bb_iqn_cfg_default(bb_iqn_cfg* c) {  // iqn/config.rs:50-67
    memset(c, 0, sizeof(*c));
    net_default(&c->f_config);
    net_default(&c->m_config);
    opt_default(&c->opt_config, 0.0);
    c->feature_dim = 64; c->embed_dim = 64;
    c->soft_update_interval = 1; c->n_updates_per_opt = 1; c->batch_size = 1; c->discount_factor = 0.99;
    c->tau = 0.005; c->train = 0; c->sample_percents_pred = BB_IQN_UNIFORM64; c->sample_percents_tgt = BB_IQN_UNIFORM64;
    c->sample_percents_act = BB_IQN_UNIFORM32; c->eps_start = 1.0; c->eps_final = 0.02; c->final_step = 100000;
    c->device = 0; c->explorer_seed = 0x0123456789abcdefULL; c->tau_seed = 0x7a07a0ULL;
}

int32_t bb_dqn_create(const bb_dqn_cfg* cfg, bb_agent** out) {
    BB_API_BEGIN
    BB_CHECK(cfg && out, "null argument");
    BB_CHECK(cfg->device >= 0, "No device is given for DQN agent");  // dqn/base.rs:258
    *out = new bb_agent{std::unique_ptr<bb::Agent>(bb::make_dqn(*cfg))};
    BB_API_END
}
int32_t bb_sac_create(const bb_sac_cfg* cfg, bb_agent** out) {
    BB_API_BEGIN
    BB_CHECK(cfg && out, "null argument");
    *out = new bb_agent{std::unique_ptr<bb::Agent>(bb::make_sac(*cfg))};
    BB_API_END
}
int32_t bb_iqn_create(const bb_iqn_cfg* cfg, bb_agent** out) {
    BB_API_BEGIN
    BB_CHECK(cfg && out, "null argument");
    *out = new bb_agent{std::unique_ptr<bb::Agent>(bb::make_iqn(*cfg))};
    BB_API_END
}
int32_t bb_agent_destroy(bb_agent* a) {
    BB_API_BEGIN
    delete a;
    BB_API_END
}
int32_t bb_agent_set_stream(bb_agent* a, void* s) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    ag.ctx.stream = s ? (cudaStream_t)s : bb::device_stream(ag.device);
    BB_API_END
}
int32_t bb_agent_set_precision(bb_agent* a, int32_t fast) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    const int passes = fast ? 1 : 3;
    ag.ctx.passes = passes;
    for (int i = 0; i < 2; ++i) ag.side_ctx[i].passes = passes;
    ag.precision_changed();   // captured CUDA graphs hold the kernels of the old mode
    BB_API_END
}
int32_t bb_agent_set_train(bb_agent* a, int32_t train) {
    BB_API_BEGIN
    A(a).train = train != 0;
    BB_API_END
}
int32_t bb_agent_is_train(const bb_agent* a, int32_t* out) {
    BB_API_BEGIN
    BB_CHECK(a && a->impl && out, "null argument");
    *out = a->impl->train ? 1 : 0;
    BB_API_END
}
int32_t bb_agent_sample(bb_agent* a, const void* obs, size_t n, void* act_out) {
    BB_API_BEGIN
    BB_CHECK(obs && act_out, "null argument");
    bb::check_device_error("bb_agent_sample");
    A(a).sample(obs, n, act_out);
    BB_API_END
}
int32_t bb_actor_step(bb_agent* a, bb_replay* rb, const void* obs, const void* reset_obs, float reward, int8_t is_terminated,
                      int8_t is_truncated, int64_t* act_out) {
    BB_API_BEGIN
    BB_CHECK(rb && obs && act_out, "null argument");
    bb::check_device_error("bb_actor_step");
    A(a).actor_step(rb->impl, obs, reset_obs, reward, is_terminated, is_truncated, act_out);
    BB_API_END
}
int32_t bb_actor_step_dev(bb_agent* a, bb_replay* rb, const void* obs_dev, const void* reset_obs_dev, float reward,
                          int8_t is_terminated, int8_t is_truncated, int64_t* act_out) {
    BB_API_BEGIN
    BB_CHECK(rb && obs_dev && act_out, "null argument");
    bb::check_device_error("bb_actor_step_dev");
    A(a).actor_step(rb->impl, obs_dev, reset_obs_dev, reward, is_terminated, is_truncated, act_out, true);
    BB_API_END
}
int32_t bb_actor_step_n(bb_agent* a, bb_replay* rb, int32_t n_envs, const void* obs, const void* reset_obs, const int8_t* reset_mask,
                        const float* reward, const int8_t* is_terminated, const int8_t* is_truncated, int64_t* act_out,
                        int32_t obs_on_device) {
    BB_API_BEGIN
    BB_CHECK(rb && obs && reward && is_terminated && is_truncated && act_out, "null argument");
    bb::check_device_error("bb_actor_step_n");
    A(a).actor_step_n(rb->impl, n_envs, obs, reset_obs, reset_mask, reward, is_terminated, is_truncated, act_out, obs_on_device != 0);
    BB_API_END
}
int32_t bb_actor_reset(bb_agent* a) {
    BB_API_BEGIN
    A(a).actor_reset();
    BB_API_END
}
int32_t bb_agent_opt(bb_agent* a, bb_replay* rb, bb_record* record) {
    BB_API_BEGIN
    BB_CHECK(rb, "null replay handle");
    // a device-side bounded wait that timed out in an EARLIER call (tcgen05 / TMA pipeline, peer barrier) left garbage in
    // the model: refuse to train on (the flag is pinned mapped memory: a plain host read, no sync)
    bb::check_device_error("bb_agent_opt");
    A(a).opt(rb->impl, record);
    BB_API_END
}
// One opt() call with a CUDA event after every kernel; writes "label ms\n" lines (label =
// phase:layer:kernel) into text_out.  Measurement aid for bench.py's roofline, not a hot path.
int32_t bb_agent_opt_profiled(bb_agent* a, bb_replay* rb, char* text_out, size_t cap) {
    BB_API_BEGIN
    BB_CHECK(rb && text_out && cap > 0, "null argument");
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    bb::Profiler prof;
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    ag.ctx.prof = &prof;
    ag.ctx.phase = "start"; ag.ctx.layer = "";
    ag.ctx.mark("begin");
    try {
        ag.opt(rb->impl, nullptr);
    } catch (...) {
        ag.ctx.prof = nullptr;
        prof.clear();
        throw;
    }
    ag.ctx.prof = nullptr;
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    std::string out;
    for (size_t i = 1; i < prof.marks.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, prof.marks[i - 1].second, prof.marks[i].second);
        char line[256];
        snprintf(line, sizeof(line), "%s %.6f\n", prof.marks[i].first.c_str(), ms);
        out += line;
    }
    prof.clear();
    BB_CHECK(out.size() + 1 <= cap, "profile text buffer too small");
    memcpy(text_out, out.c_str(), out.size() + 1);
    BB_API_END
}
int32_t bb_agent_n_opts(const bb_agent* a, uint64_t* out) {
    BB_API_BEGIN
    BB_CHECK(a && a->impl && out, "null argument");
    *out = a->impl->n_opts;
    BB_API_END
}
int32_t bb_agent_save_params(bb_agent* a, const char* dir) {
    BB_API_BEGIN
    BB_CHECK(dir, "null path");
    A(a).save_params(dir);
    BB_API_END
}
int32_t bb_agent_load_params(bb_agent* a, const char* dir) {
    BB_API_BEGIN
    BB_CHECK(dir, "null path");
    A(a).load_params(dir);
    BB_API_END
}
int32_t bb_agent_param_count(bb_agent* a, const char* model, uint64_t* n_tensors, uint64_t* n_floats) {
    BB_API_BEGIN
    bb::Model* m = A(a).model(model);
    uint64_t nf = 0;
    for (auto& pi : m->params) nf += pi.numel;
    if (n_tensors) *n_tensors = m->params.size();
    if (n_floats) *n_floats = nf;
    BB_API_END
}
int32_t bb_agent_param_info(bb_agent* a, const char* model, uint64_t index, char* name_out, size_t name_cap,
                            int64_t* shape_out, int32_t* ndim_out) {
    BB_API_BEGIN
    bb::Model* m = A(a).model(model);
    BB_CHECK(index < m->params.size(), "parameter index out of range");
    const bb::ParamInfo& pi = m->params[index];
    if (name_out && name_cap) {
        strncpy(name_out, pi.name.c_str(), name_cap - 1);
        name_out[name_cap - 1] = 0;
    }
    if (shape_out)
        for (size_t i = 0; i < 4; ++i) shape_out[i] = i < pi.shape.size() ? pi.shape[i] : 1;
    if (ndim_out) *ndim_out = (int32_t)pi.shape.size();
    BB_API_END
}

static void get_tensor(bb::Agent& ag, const float* dev_base, const bb::ParamInfo& pi, float* host_out) {
    std::vector<float> tmp(pi.numel);
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    BB_CUDA(cudaMemcpy(tmp.data(), dev_base + pi.offset, pi.numel * 4, cudaMemcpyDeviceToHost));
    bb::param_to_reference(pi, tmp.data(), host_out);
}

int32_t bb_agent_get_param(bb_agent* a, const char* model, const char* name, float* host_out, size_t n) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    bb::Model* m = ag.model(model);
    const bb::ParamInfo* pi = m->find(name);
    BB_CHECK(pi, "no such parameter");
    BB_CHECK(n == pi->numel, "parameter size mismatch");
    get_tensor(ag, m->p, *pi, host_out);
    BB_API_END
}
int32_t bb_agent_set_param(bb_agent* a, const char* model, const char* name, const float* host_in, size_t n) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    bb::Model* m = ag.model(model);
    const bb::ParamInfo* pi = m->find(name);
    BB_CHECK(pi, "no such parameter");
    BB_CHECK(n == pi->numel, "parameter size mismatch");
    std::vector<float> tmp(pi->numel);
    bb::param_to_internal(*pi, host_in, tmp.data());
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    bb::h2d_sync(m->p + pi->offset, tmp.data(), pi->numel * 4, ag.ctx.stream);
    bb::make_lo(ag.ctx, m->p + pi->offset, m->p_lo() + pi->offset, pi->numel);
    BB_API_END
}
int32_t bb_agent_reset_opt_state(bb_agent* a) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    for (bb::Model* m : ag.models)
        if (m->has_opt) {
            BB_CUDA(cudaMemsetAsync(m->m, 0, m->n * 4, ag.ctx.stream));
            BB_CUDA(cudaMemsetAsync(m->v, 0, m->n * 4, ag.ctx.stream));
            m->step = 0;
        }
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    BB_API_END
}
int32_t bb_agent_get_opt_state(bb_agent* a, const char* model, const char* name, float* host_m, float* host_v, size_t n,
                               uint64_t* step) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    bb::Model* m = ag.model(model);
    BB_CHECK(m->has_opt, "this VarStore has no optimizer state");
    const bb::ParamInfo* pi = m->find(name);
    BB_CHECK(pi, "no such parameter");
    BB_CHECK(n == pi->numel, "parameter size mismatch");
    if (host_m) get_tensor(ag, m->m, *pi, host_m);
    if (host_v) get_tensor(ag, m->v, *pi, host_v);
    if (step) *step = m->step;
    BB_API_END
}

// SyncModel: NamedTensors::copy_from(var_store) (util/named_tensors.rs:11-36) as one flat blob in
// reference layout, tensors in VarStore order.
int32_t bb_agent_model_info_size(bb_agent* a, uint64_t* n_floats) {
    BB_API_BEGIN
    bb::Model* m = A(a).sync_model_src();
    uint64_t nf = 0;
    for (auto& pi : m->params) nf += pi.numel;
    *n_floats = nf;
    BB_API_END
}
int32_t bb_agent_model_info(bb_agent* a, float* host_out, size_t n, uint64_t* n_opts) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    bb::Model* m = ag.sync_model_src();
    std::vector<float> h(m->n);
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    BB_CUDA(cudaMemcpy(h.data(), m->p, m->n * 4, cudaMemcpyDeviceToHost));
    size_t off = 0;
    for (auto& pi : m->params) {
        BB_CHECK(off + pi.numel <= n, "model_info buffer too small");
        bb::param_to_reference(pi, h.data() + pi.offset, host_out + off);
        off += pi.numel;
    }
    if (n_opts) *n_opts = ag.n_opts;
    BB_API_END
}
int32_t bb_agent_sync_model(bb_agent* a, const float* host_in, size_t n) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    bb::Model* m = ag.sync_model_src();
    std::vector<float> h(m->n, 0.f);
    size_t off = 0;
    for (auto& pi : m->params) {
        BB_CHECK(off + pi.numel <= n, "model_info blob too small");
        bb::param_to_internal(pi, host_in + off, h.data() + pi.offset);
        off += pi.numel;
    }
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    bb::h2d_sync(m->p, h.data(), m->n * 4, ag.ctx.stream);
    m->refresh_lo(ag.ctx);
    BB_API_END
}
int32_t bb_agent_sync_model_from(bb_agent* dst, bb_agent* src) {
    BB_API_BEGIN
    bb::Agent &d = A(dst), &s = A(src);
    BB_CHECK(d.device == s.device, "sync_model_from needs both agents on one GPU");
    bb::DeviceGuard g(d.device);
    bb::Model *md = d.sync_model_src(), *ms = s.sync_model_src();
    if (d.ctx.stream != s.ctx.stream) bb::stream_wait(d.ctx.stream, s.ctx.stream);
    md->copy_params_from(*ms, d.ctx.stream);
    BB_API_END
}
int32_t bb_agent_inject_noise(bb_agent* a, int32_t slot, const float* host, size_t n) {
    BB_API_BEGIN
    BB_CHECK(host, "null argument");
    A(a).inject_noise(slot, host, n);
    BB_API_END
}
int32_t bb_agent_exchange_trace(bb_agent* a, uint32_t* out32) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    BB_CUDA(cudaStreamSynchronize(ag.ctx.stream));
    BB_CUDA(cudaMemcpy(out32, ag.xchg_ctr, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    BB_API_END
}
int32_t bb_agent_grad_buffer(bb_agent* a, void** dev_ptr, uint64_t* n_floats) {
    BB_API_BEGIN
    BB_CHECK(dev_ptr && n_floats, "null argument");
    A(a).grad_buffer(dev_ptr, n_floats);
    BB_API_END
}
int32_t bb_agent_ipc_export(bb_agent* a, void* handle_out, void* flag_handle_out) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    void* gp;
    uint64_t n;
    ag.grad_buffer(&gp, &n);
    BB_CHECK(gp, "agent has no gradient buffer to share");
    if (!ag.my_flags) ag.my_flags = bb::dev_alloc_zero<unsigned int>(64);
    BB_CUDA(cudaDeviceSynchronize());
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    BB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle_out, gp));
    BB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)flag_handle_out, ag.my_flags));
    BB_API_END
}
int32_t bb_agent_ipc_connect(bb_agent* a, int32_t rank, int32_t world, const void* handles, const void* flag_handles) {
    BB_API_BEGIN
    bb::Agent& ag = A(a);
    bb::DeviceGuard g(ag.device);
    BB_CHECK(world >= 1 && world <= 8 && rank >= 0 && rank < world, "bad rank/world");
    void* gp;
    uint64_t n;
    ag.grad_buffer(&gp, &n);
    BB_CHECK(ag.my_flags, "call bb_agent_ipc_export first");
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            ag.peer_grad[r] = (const float*)gp;
            ag.peer_flag[r] = ag.my_flags;
            continue;
        }
        cudaIpcMemHandle_t h;
        void* p = nullptr;
        memcpy(&h, (const char*)handles + 64 * r, 64);
        BB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ag.peer_grad[r] = (const float*)p;
        memcpy(&h, (const char*)flag_handles + 64 * r, 64);
        BB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ag.peer_flag[r] = (unsigned int*)p;
    }
    ag.rank = rank;
    ag.world = world;
    BB_API_END
}

}  // extern "C"
