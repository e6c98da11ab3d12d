// sac.cu -- Sac agent (border-tch-agent/src/sac/base.rs) on the device.
//
//   action_logp    sac/base.rs:73-87      qvals_min      sac/base.rs:97-105
//   update_critic  sac/base.rs:107-149    update_actor   sac/base.rs:151-167
//   soft_update    sac/base.rs:169-173    opt_           sac/base.rs:175-198
//   EntCoef        sac/ent_coef.rs:27-75  Mlp2 (actor)   mlp/mlp2.rs:23-50
//   Critic = Mlp as SubModel2: cat[obs, act] -> mlp.ln*  mlp/base.rs:83-88
//
// The reference relies on libtorch autograd; the backward passes are written out here:
//   a = tanh(u), u = std*z + mean, std = exp(clip(s, lo, hi)), s = exp(sl(x))   (double exp, as
//   in the reference: Mlp2 already returns exp(head2) and action_logp exponentiates again)
//   logp = sum_j(-0.5 ln 2pi - 0.5 z_j^2) - sum_j ln(1 - a_j^2 + eps)
//   actor loss = mean(alpha*logp - min_i Q_i(o, a))
#include <math.h>
#include "agent.cuh"

namespace bb {

__device__ __forceinline__ unsigned long long sac_mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// z ~ N(0,1): counter-based Box-Muller (the reference draws Tensor::randn on the CPU generator;
// parity tests inject z instead)
__device__ __forceinline__ float sac_normal(unsigned long long seed, unsigned long long ctr) {
    unsigned long long h = sac_mix64(seed ^ (ctr * 0xD1342543DE82EF95ull));
    float u1 = ((float)(uint32_t)(h >> 40) + 1.0f) * (1.0f / 16777216.0f);
    float u2 = (float)(uint32_t)((h >> 8) & 0xffffff) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// heads [B][2A] = (mean | h2) -> a [B][A], logp [B]; also writes XA[b] = [obs_b, a_b] for the critics
__global__ void sac_action_kernel(const float* __restrict__ heads, const float* __restrict__ z_in, float* __restrict__ z_out,
                                  const float* __restrict__ obs, int obs_dim, float* __restrict__ xa,
                                  float* __restrict__ a_out, float* __restrict__ logp, int B, int A, float lo, float hi,
                                  float eps, unsigned long long seed, const unsigned long long* __restrict__ ctr_dev) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const unsigned long long ctr0 = *ctr_dev;   // (device memory: the launch is argument-invariant, see SacScalars)
    const float* hd = heads + (size_t)b * 2 * A;
    float lp_n = 0.f, lp_t = 0.f;
    for (int j = 0; j < A; ++j) {
        float mean = hd[j];
        float s = expf(hd[A + j]);                  // Mlp2: exp(head2)
        float sd = expf(fminf(fmaxf(s, lo), hi));   // lstd.clip(min, max).exp()
        float z = z_in ? z_in[(size_t)b * A + j] : sac_normal(seed, ctr0 + (unsigned long long)b * A + j);
        if (z_out) z_out[(size_t)b * A + j] = z;
        float a = tanhf(sd * z + mean);
        a_out[(size_t)b * A + j] = a;
        xa[(size_t)b * (obs_dim + A) + obs_dim + j] = a;
        lp_n += -0.5f * 1.8378770664093453f - 0.5f * z * z;  // normal_logp (sac/base.rs:25-29)
        lp_t += logf(1.0f - a * a + eps);
    }
    for (int j = 0; j < obs_dim; ++j) xa[(size_t)b * (obs_dim + A) + j] = obs[(size_t)b * obs_dim + j];
    logp[b] = lp_n - lp_t;
}

__global__ void concat2_kernel(const float* __restrict__ x, int dx, const float* __restrict__ y, int dy,
                               float* __restrict__ out, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int d = dx + dy;
    if (i >= B * d) return;
    int b = i / d, j = i % d;
    out[i] = j < dx ? x[(size_t)b * dx + j] : y[(size_t)b * dy + (j - dx)];
}

__device__ __forceinline__ float sac_block_sum(float v, float* s) {
    __syncthreads();
    s[threadIdx.x] = v;
    __syncthreads();
    for (int k = blockDim.x / 2; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
        __syncthreads();
    }
    return s[0];
}

// EntCoef::update (ent_coef.rs:69-75): loss = -mean(log_alpha * (logp + target)); Adam on the scalar.
__global__ void __launch_bounds__(1024) sac_entcoef_kernel(const float* __restrict__ logp, int B, float target,
                                                           float* log_alpha, float* m_, float* v_, float lr,
                                                           const float* __restrict__ bc /* bc1, sqrt(bc2) */) {
    __shared__ float s[1024];
    const float bc1 = bc[0], bc2_sqrt = bc[1];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) acc += logp[b] + target;
    float tot = sac_block_sum(acc, s);
    if (threadIdx.x == 0) {
        float g = -(tot / (float)B);
        float m = m_[0] * 0.9f + 0.1f * g;
        float v = v_[0] * 0.999f + 0.001f * g * g;
        float denom = sqrtf(v) / bc2_sqrt + 1e-8f;
        log_alpha[0] = log_alpha[0] + (-(lr / bc1)) * (m / denom);
        m_[0] = m; v_[0] = v;
    }
}

struct QPtrs { const float* q[8]; float* dq[8]; };

// Everything that changes from one update to the next, in device memory: one tiny launch writes it (kernel parameters are
// copied at launch time, so the host can run ahead), every other launch of the update reads it and is argument-invariant.
struct SacScalars {
    AdamScalars pi, q[8];
    float ent_bc1, ent_bc2_sqrt;
    unsigned long long ctr[2];   // noise counters of update_actor's and update_critic's action_logp
};
__global__ void sac_set_scalars_kernel(SacScalars s, SacScalars* dst) {
    if (threadIdx.x == 0) *dst = s;
}

// actor loss pieces (sac/base.rs:151-164): qmin over critics, dL/dQ_i, loss = mean(alpha*logp - qmin)
__global__ void __launch_bounds__(1024) sac_actor_loss_kernel(QPtrs qp, int n_critics, const float* __restrict__ logp,
                                                              const float* __restrict__ ent_state, int B, float* out) {
    __shared__ float s[1024];
    const float alpha = expf(ent_state[0]);
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int best = 0;
        float qm = qp.q[0][b];
        for (int i = 1; i < n_critics; ++i)
            if (qp.q[i][b] < qm) { qm = qp.q[i][b]; best = i; }
        for (int i = 0; i < n_critics; ++i) qp.dq[i][b] = (i == best) ? -1.0f / (float)B : 0.f;
        acc += alpha * logp[b] - qm;
    }
    float tot = sac_block_sum(acc, s);
    if (threadIdx.x == 0) { out[1] = tot / (float)B; out[2] = alpha; }
}

// dL/d(heads) from dL/da (through the critics) and dL/dlogp = alpha/B
__global__ void sac_actor_grad_kernel(const float* __restrict__ heads, const float* __restrict__ z,
                                      const float* __restrict__ a, QPtrs dxa, int n_critics, int obs_dim,
                                      const float* __restrict__ ent_state, float* __restrict__ dheads, int B, int A,
                                      float lo, float hi, float eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * A) return;
    int b = i / A, j = i % A;
    const float alpha = expf(ent_state[0]);
    float av = a[i];
    float dq = 0.f;
    for (int c = 0; c < n_critics; ++c) dq += dxa.q[c][(size_t)b * (obs_dim + A) + obs_dim + j];
    float one_m = 1.0f - av * av;
    float ga = (alpha / (float)B) * (2.0f * av / (one_m + eps)) + dq;
    float gu = ga * one_m;
    float s = expf(heads[(size_t)b * 2 * A + A + j]);
    float gate = (s >= lo && s <= hi) ? 1.f : 0.f;
    float sd = expf(fminf(fmaxf(s, lo), hi));
    dheads[(size_t)b * 2 * A + j] = gu;
    dheads[(size_t)b * 2 * A + A + j] = gu * z[i] * sd * gate * s;
}

// critic target and per-critic loss gradients (sac/base.rs:107-135)
__global__ void __launch_bounds__(1024) sac_critic_loss_kernel(QPtrs pred, QPtrs tgtq, int n_critics,
                                                               const float* __restrict__ next_logp,
                                                               const float* __restrict__ reward, const int8_t* __restrict__ term,
                                                               const float* __restrict__ ent_state, int B, float gamma,
                                                               float reward_scale, int loss_kind, float* out) {
    __shared__ float s[1024];
    const float alpha = expf(ent_state[0]);
    float acc = 0.f;
    const float invB = 1.0f / (float)B;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float qm = tgtq.q[0][b];
        for (int i = 1; i < n_critics; ++i) qm = fminf(qm, tgtq.q[i][b]);
        float next_q = __fsub_rn(qm, __fmul_rn(alpha, next_logp[b]));
        float nt = __fsub_rn(1.0f, (float)term[b]);
        float tgt = __fadd_rn(__fmul_rn(reward_scale, reward[b]), __fmul_rn(__fmul_rn(nt, gamma), next_q));
        for (int i = 0; i < n_critics; ++i) {
            float d = pred.q[i][b] - tgt;
            float l, g;
            if (loss_kind == BB_LOSS_SMOOTH_L1) {
                float ad = fabsf(d);
                l = ad < 1.f ? 0.5f * d * d : ad - 0.5f;
                g = ad < 1.f ? d : (d > 0.f ? 1.f : -1.f);
            } else {
                l = d * d;
                g = 2.f * d;
            }
            pred.dq[i][b] = g * invB;
            acc += l;
        }
    }
    float tot = sac_block_sum(acc, s);
    if (threadIdx.x == 0) out[0] = tot * invB / (float)n_critics;  // mean over critics of the mean losses
}

struct Sac : Agent {
    bb_sac_cfg cfg;
    int obs_dim, act_dim, n_critics;
    Net pi_net, q_net;
    Model pi, ent;
    std::vector<Model> qnets, qnets_tgt;
    NetWorkspace ws_pi, ws_pi_next, ws_pi_act;
    std::vector<NetWorkspace> ws_qa, ws_qc;
    NetWorkspace ws_qt;
    int ws_batch = 0;
    float *d_a = nullptr, *d_a_next = nullptr, *d_logp = nullptr, *d_logp_next = nullptr, *d_z = nullptr,
          *d_xa = nullptr, *d_xa2 = nullptr, *d_xa_c = nullptr, *d_dxa = nullptr, *d_qt = nullptr, *d_out = nullptr;
    float* d_inject[2] = {nullptr, nullptr};
    size_t inject_n[2] = {0, 0};
    uint64_t noise_ctr = 0;
    SacScalars* d_sc = nullptr;
    UpdateGraph graph;
    FastRand fr;
    float *h_in = nullptr, *h_heads = nullptr;
    float* d_in = nullptr;
    size_t in_cap = 0;

    explicit Sac(const bb_sac_cfg& c) : cfg(c), fr(c.noise_seed) {
        init_base(c.device);
        DeviceGuard g(device);
        train = c.train != 0;
        n_critics = (int)c.n_critics;
        BB_CHECK(n_critics >= 1 && n_critics <= 8, "n_critics must be in 1..8");
        BB_CHECK(c.pi_config.kind == BB_NET_MLP && c.q_config.kind == BB_NET_MLP, "SAC nets are Mlp2 / Mlp");
        BB_CHECK(c.pi_config.n_units >= 1, "Mlp2 needs at least one hidden layer (mlp2.rs:33)");
        obs_dim = c.pi_config.in_dim;
        act_dim = c.pi_config.out_dim;
        BB_CHECK(c.q_config.in_dim == obs_dim + act_dim, "critic in_dim must be obs_dim + act_dim");
        // actor: Mlp2 (mlp2.rs:31-50): trunk mlp.al{i} + ReLU, heads ml / sl
        pi_net.reset();
        int in = obs_dim;
        for (int i = 0; i < c.pi_config.n_units; ++i) {
            pi_net.add_linear_layer("mlp.al" + std::to_string(i), in, c.pi_config.units[i], true);
            in = c.pi_config.units[i];
        }
        pi_net.add_twin_heads("ml", "sl", in, act_dim);
        pi_net.in_elems = obs_dim;
        // critic: Mlp as SubModel2 (mlp/base.rs:83-106): hidden mlp.ln{i}+ReLU, final mlp.ln{n}; no activation_out
        bb_net_cfg qc = c.q_config;
        qc.activation_out = 0;
        q_net.build(qc, "");
        pi.name = "pi"; pi.params = pi_net.params; pi.n = pi_net.n_params; pi.alloc(true); pi.set_hyper(c.pi_opt_config);
        pi_net.init_params(ctx, pi.p, c.init_seed * 7919 + 1);
        qnets.resize(n_critics); qnets_tgt.resize(n_critics);
        for (int i = 0; i < n_critics; ++i) {
            Model& q = qnets[i];
            q.name = "qnet_" + std::to_string(i); q.params = q_net.params; q.n = q_net.n_params; q.alloc(true);
            q.set_hyper(c.q_opt_config);
            q_net.init_params(ctx, q.p, c.init_seed * 7919 + 100 + i);
            Model& t = qnets_tgt[i];
            t.name = "qnet_tgt_" + std::to_string(i); t.params = q_net.params; t.n = q_net.n_params; t.alloc(false);
            t.copy_params_from(q, ctx.stream);  // qnets.push(critic.clone()); qnets_tgt.push(critic)
        }
        // EntCoef: VarStore with one variable "log_alpha" [1] (ent_coef.rs:27-48)
        ent.name = "ent_coef";
        ParamInfo pa; pa.name = "log_alpha"; pa.shape = {1}; pa.offset = 0; pa.numel = 1; pa.perm = 0; pa.pc = pa.ph = pa.pw = 0; pa.fan_in = 1;
        ent.params = {pa}; ent.n = 4;
        ent.alloc(true);
        float la = c.ent_coef_mode == BB_ENTCOEF_FIX ? (float)log(c.ent_coef_fix) : 0.0f;
        BB_CUDA(cudaMemcpyAsync(ent.p, &la, 4, cudaMemcpyHostToDevice, ctx.stream));
        ent.hyper = AdamHyper{c.ent_coef_lr, 0.9, 0.999, 1e-8, 0.0, false};
        models.push_back(&pi);
        models.push_back(&ent);
        for (auto& q : qnets) models.push_back(&q);
        for (auto& q : qnets_tgt) models.push_back(&q);
        d_out = dev_alloc_zero<float>(8, ctx.stream);
        d_sc = dev_alloc_zero<SacScalars>(1, ctx.stream);
        BB_CUDA(cudaStreamSynchronize(ctx.stream));
        ws_qa.resize(n_critics); ws_qc.resize(n_critics);
        pi_net.alloc_workspace(ws_pi_act, 1, false);
    }

    ~Sac() override {
        DeviceGuard g(device);
        cudaStreamSynchronize(ctx.stream);
        graph.reset();
        cudaFree(d_sc);
        ws_pi.release(); ws_pi_next.release(); ws_pi_act.release(); ws_qt.release();
        for (auto& w : ws_qa) w.release();
        for (auto& w : ws_qc) w.release();
        pi.release(); ent.release();
        for (auto& q : qnets) q.release();
        for (auto& q : qnets_tgt) q.release();
        cudaFree(d_a); cudaFree(d_a_next); cudaFree(d_logp); cudaFree(d_logp_next); cudaFree(d_z); cudaFree(d_xa); cudaFree(d_xa_c);
        cudaFree(d_xa2); cudaFree(d_dxa); cudaFree(d_qt); cudaFree(d_out); cudaFree(d_inject[0]);
        cudaFree(d_inject[1]); cudaFree(d_in);
        if (h_in) cudaFreeHost(h_in);
        if (h_heads) cudaFreeHost(h_heads);
    }
    Model* sync_model_src() override { return &pi; }  // sac/base.rs:377-386 ships only pi
    void precision_changed() override { graph.reset(); }
    void grad_buffer(void** p, uint64_t* n) override { *p = pi.g; *n = pi.n; }

    void inject_noise(int slot, const float* host, size_t n) override {
        DeviceGuard g(device);
        BB_CHECK(slot == 0 || slot == 1, "SAC noise slots: 0 = update_actor z, 1 = update_critic z'");
        BB_CUDA(cudaStreamSynchronize(ctx.stream));
        cudaFree(d_inject[slot]);
        d_inject[slot] = dev_alloc<float>(n);
        h2d_sync(d_inject[slot], host, n * 4, ctx.stream);
        inject_n[slot] = n;
    }

    void ensure_ws(int B) {
        if (B <= ws_batch) return;
        BB_CUDA(cudaStreamSynchronize(ctx.stream));
        pi_net.alloc_workspace(ws_pi, B, true);
        pi_net.alloc_workspace(ws_pi_next, B, false);
        for (int i = 0; i < n_critics; ++i) {
            q_net.alloc_workspace(ws_qa[i], B, true);
            q_net.alloc_workspace(ws_qc[i], B, true);
        }
        q_net.alloc_workspace(ws_qt, B, false);
        for (float** p : {&d_a, &d_a_next, &d_z}) { cudaFree(*p); *p = dev_alloc<float>((size_t)B * act_dim); }
        for (float** p : {&d_logp, &d_logp_next}) { cudaFree(*p); *p = dev_alloc<float>(B); }
        for (float** p : {&d_xa, &d_xa2, &d_xa_c}) { cudaFree(*p); *p = dev_alloc<float>((size_t)B * (obs_dim + act_dim)); }
        cudaFree(d_dxa); d_dxa = dev_alloc<float>((size_t)n_critics * B * (obs_dim + act_dim));
        cudaFree(d_qt); d_qt = dev_alloc<float>((size_t)n_critics * B);
        ws_batch = B;
    }

    // action_logp (sac/base.rs:73-87) on `obs`; leaves [obs, a] in xa
    void action_logp(const float* obs, int B, NetWorkspace& w, int slot, float* a, float* logp, float* z_keep, float* xa) {
        const float* heads = pi_net.forward(ctx, pi.p, obs, obs_dim, B, w);
        const float* zin = nullptr;
        if (inject_n[slot]) {
            BB_CHECK(inject_n[slot] == (size_t)B * act_dim, "injected noise has the wrong size");
            zin = d_inject[slot];
        }
        sac_action_kernel<<<(B + 127) / 128, 128, 0, ctx.stream>>>(heads, zin, z_keep, obs, obs_dim, xa, a, logp, B, act_dim,
                                                                    (float)cfg.min_lstd, (float)cfg.max_lstd, (float)cfg.epsilon,
                                                                    cfg.noise_seed, &d_sc->ctr[slot]);
        BB_LAUNCHED();
        ctx.layer = "pi"; ctx.mark("sac_action");
    }

    QPtrs qptrs(std::vector<NetWorkspace>& ws) {
        QPtrs p{};
        for (int i = 0; i < n_critics; ++i) { p.q[i] = ws[i].act.back(); p.dq[i] = ws[i].dact.back(); }
        return p;
    }

    void update(Replay& rb, float* h_rec /* loss_critic, loss_actor, alpha */) {
        const int B = (int)cfg.batch_size;
        BB_CHECK(B >= 1 && B <= 65536, "batch_size out of range");
        ensure_ws(B);
        BB_CHECK(rb.cfg.obs_kind == BB_F32 && (int)rb.cfg.obs_elems == obs_dim, "replay obs rows must be f32[obs_dim]");
        BB_CHECK(rb.cfg.act_kind == BB_F32 && (int)rb.cfg.act_elems == act_dim, "replay act rows must be f32[act_dim]");
        // host-side bookkeeping of this update: optimizer steps, noise counters -> SacScalars (one launch)
        SacScalars sc{};
        const bool auto_ent = cfg.ent_coef_mode == BB_ENTCOEF_AUTO;
        if (auto_ent) {
            ent.step += 1;
            sc.ent_bc1 = (float)(1.0 - pow(0.9, (double)ent.step));
            sc.ent_bc2_sqrt = (float)sqrt(1.0 - pow(0.999, (double)ent.step));
        }
        pi.step += 1;
        sc.pi = adam_scalars(pi.hyper, pi.step);
        for (int i = 0; i < n_critics; ++i) { qnets[i].step += 1; sc.q[i] = adam_scalars(qnets[i].hyper, qnets[i].step); }
        sc.ctr[0] = noise_ctr; sc.ctr[1] = noise_ctr + (uint64_t)B * act_dim;
        noise_ctr += 2 * (uint64_t)B * act_dim;
        sac_set_scalars_kernel<<<1, 32, 0, ctx.stream>>>(sc, d_sc);
        BB_LAUNCHED();
        ctx.phase = "replay"; ctx.layer = "scalars"; ctx.mark("sac_set_scalars");
        bb_batch_view bv;
        bool launch_sample = true;
        auto enqueue = [&]() { enqueue_update(rb, B, bv, launch_sample); };
        graph.run(ctx, rb, B, inject_n[0] == 0 && inject_n[1] == 0, enqueue, [&]() { rb.sample(B, &bv, false); });
        n_opts += 1;
        inject_n[0] = inject_n[1] = 0;
        if (h_rec) {
            BB_CUDA(cudaMemcpyAsync(h_scratch, d_out, 8 * sizeof(float), cudaMemcpyDeviceToHost, ctx.stream));
            BB_CUDA(cudaStreamSynchronize(ctx.stream));
            h_rec[0] += h_scratch[0]; h_rec[1] += h_scratch[1]; h_rec[2] = h_scratch[2];
        }
    }

    // every launch of one update (sac/base.rs:175-198 order: update_actor, update_critic, soft_update); argument-invariant
    void enqueue_update(Replay& rb, int B, bb_batch_view& bv, bool launch_sample) {
        const bool auto_ent = cfg.ent_coef_mode == BB_ENTCOEF_AUTO;
        if (rb.stream != ctx.stream) stream_wait(rb.stream, ctx.stream);
        rb.sample(B, &bv, launch_sample);  // buffer.batch(self.batch_size), sac/base.rs:180
        if (rb.stream != ctx.stream) stream_wait(ctx.stream, rb.stream);
        ctx.phase = "replay"; ctx.layer = "batch"; ctx.mark("sample_gather");
        const float* obs = (const float*)bv.obs;
        const float* next_obs = (const float*)bv.next_obs;
        const int D = obs_dim + act_dim;

        // The critics' forward on the replayed actions (first half of update_critic, sac/base.rs:107-118) does not depend on
        // the actor update -- the critics' parameters only change at the end of the step -- so it runs beside it on a side
        // stream, into its own [obs, act] buffer.
        const bool conc = ctx.concurrent();
        const Ctx& cctx = conc ? *ctx.side[0] : ctx;
        auto critic_forward = [&]() {
            cctx.phase = "critic";
            concat2_kernel<<<(B * D + 255) / 256, 256, 0, cctx.stream>>>(obs, obs_dim, (const float*)bv.act, act_dim, d_xa_c, B);
            BB_LAUNCHED();
            cctx.layer = "cat"; cctx.mark("concat");
            for (int i = 0; i < n_critics; ++i) q_net.forward(cctx, qnets[i].p, d_xa_c, D, B, ws_qc[i]);
        };
        if (conc) { ctx.fork_to(cctx); critic_forward(); }

        // ---------------- update_actor (sac/base.rs:151-167)
        ctx.phase = "actor";
        action_logp(obs, B, ws_pi, 0, d_a, d_logp, d_z, d_xa);
        if (auto_ent) {  // ent_coef.update(&log_p.detach())
            sac_entcoef_kernel<<<1, 1024, 0, ctx.stream>>>(d_logp, B, (float)cfg.ent_coef_target, ent.p, ent.m, ent.v,
                                                          (float)cfg.ent_coef_lr, &d_sc->ent_bc1);
            BB_LAUNCHED();
            ctx.layer = "ent_coef"; ctx.mark("entcoef_adam");
        }
        for (int i = 0; i < n_critics; ++i) q_net.forward(ctx, qnets[i].p, d_xa, D, B, ws_qa[i]);
        sac_actor_loss_kernel<<<1, 1024, 0, ctx.stream>>>(qptrs(ws_qa), n_critics, d_logp, ent.p, B, d_out);
        BB_LAUNCHED();
        ctx.layer = "loss"; ctx.mark("sac_actor_loss");
        QPtrs dxa{};
        for (int i = 0; i < n_critics; ++i) {  // dQ/d[obs,a]: data gradient only
            float* dst = d_dxa + (size_t)i * B * D;
            q_net.backward(ctx, qnets[i].p, nullptr, d_xa, D, B, ws_qa[i], dst, D);
            dxa.q[i] = dst;
        }
        sac_actor_grad_kernel<<<(B * act_dim + 127) / 128, 128, 0, ctx.stream>>>(
            ws_pi.act.back(), d_z, d_a, dxa, n_critics, obs_dim, ent.p, ws_pi.dact.back(), B, act_dim, (float)cfg.min_lstd,
            (float)cfg.max_lstd, (float)cfg.epsilon);
        BB_LAUNCHED();
        ctx.layer = "pi"; ctx.mark("sac_actor_grad");
        pi_net.backward(ctx, pi.p, pi.g, obs, obs_dim, B, ws_pi, nullptr, 0);
        adam_step(ctx, pi.p, pi.g, pi.m, pi.v, pi.n, pi.hyper, pi.step, nullptr, 1, nullptr, nullptr, 0, &d_sc->pi);

        // ---------------- update_critic (sac/base.rs:107-149)
        ctx.phase = "critic";
        if (!conc) critic_forward();
        action_logp(next_obs, B, ws_pi_next, 1, d_a_next, d_logp_next, nullptr, d_xa2);  // with the updated pi
        QPtrs tq{};
        for (int i = 0; i < n_critics; ++i) {
            const float* q = q_net.forward(ctx, qnets_tgt[i].p, d_xa2, D, B, ws_qt);
            BB_CUDA(cudaMemcpyAsync(d_qt + (size_t)i * B, q, (size_t)B * 4, cudaMemcpyDeviceToDevice, ctx.stream));
            tq.q[i] = d_qt + (size_t)i * B;
        }
        if (conc) ctx.join_from(cctx);
        sac_critic_loss_kernel<<<1, 1024, 0, ctx.stream>>>(qptrs(ws_qc), tq, n_critics, d_logp_next, bv.reward,
                                                          bv.is_terminated, ent.p, B, (float)cfg.gamma,
                                                          (float)cfg.reward_scale, cfg.critic_loss, d_out);
        BB_LAUNCHED();
        ctx.layer = "loss"; ctx.mark("sac_critic_loss");
        for (int i = 0; i < n_critics; ++i) {  // separate Adam per critic (sac/base.rs:137-139)
            q_net.backward(ctx, qnets[i].p, qnets[i].g, d_xa_c, D, B, ws_qc[i], nullptr, 0);
            adam_step(ctx, qnets[i].p, qnets[i].g, qnets[i].m, qnets[i].v, qnets[i].n, qnets[i].hyper, qnets[i].step, nullptr, 1,
                      nullptr, nullptr, 0, &d_sc->q[i]);
        }
        // ---------------- soft_update (sac/base.rs:169-173)
        ctx.phase = "target_update";
        for (int i = 0; i < n_critics; ++i) track(ctx, qnets_tgt[i].p, qnets[i].p, qnets[i].n, cfg.tau);
    }

    void opt(Replay& rb, bb_record* rec) override {  // opt_, sac/base.rs:175-198
        DeviceGuard g(device);
        float acc[3] = {0.f, 0.f, 0.f};
        for (uint64_t i = 0; i < cfg.n_updates_per_opt; ++i) update(rb, rec ? acc : nullptr);
        if (rec) {
            memset(rec, 0, sizeof(*rec));
            rec->loss_critic = acc[0] / (float)cfg.n_updates_per_opt;
            rec->loss_actor = acc[1] / (float)cfg.n_updates_per_opt;
            rec->ent_coef = acc[2];
            rec->n_opts = n_opts;
        }
    }

    // Policy::sample (sac/base.rs:216-227): tanh(std * randn + mean) in training, tanh(mean) in eval
    void sample(const void* obs, size_t n, void* act_out) override {
        DeviceGuard g(device);
        BB_CHECK(n >= 1 && n <= 4096, "sample: n out of range");
        if (n * obs_dim > in_cap || (int)n > ws_pi_act.max_batch) {
            BB_CUDA(cudaStreamSynchronize(ctx.stream));
            cudaFree(d_in);
            if (h_in) cudaFreeHost(h_in);
            if (h_heads) cudaFreeHost(h_heads);
            in_cap = n * obs_dim;
            d_in = dev_alloc<float>(in_cap);
            BB_CUDA(cudaMallocHost(&h_in, in_cap * 4));
            BB_CUDA(cudaMallocHost(&h_heads, n * 2 * act_dim * 4));
            pi_net.alloc_workspace(ws_pi_act, (int)n, false);
        }
        memcpy(h_in, obs, n * obs_dim * 4);
        BB_CUDA(cudaMemcpyAsync(d_in, h_in, n * obs_dim * 4, cudaMemcpyHostToDevice, ctx.stream));
        const float* heads = pi_net.forward(ctx, pi.p, d_in, obs_dim, (int)n, ws_pi_act);
        BB_CUDA(cudaMemcpyAsync(h_heads, heads, n * 2 * act_dim * 4, cudaMemcpyDeviceToHost, ctx.stream));
        BB_CUDA(cudaStreamSynchronize(ctx.stream));
        float* out = (float*)act_out;
        for (size_t b = 0; b < n; ++b)
            for (int j = 0; j < act_dim; ++j) {
                float mean = h_heads[b * 2 * act_dim + j];
                float v = mean;
                if (train) {
                    float s = expf(h_heads[b * 2 * act_dim + act_dim + j]);
                    float sd = expf(std::min(std::max(s, (float)cfg.min_lstd), (float)cfg.max_lstd));
                    double u1 = 1.0 - fr.f64(), u2 = fr.f64();
                    float z = (float)(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
                    v = sd * z + mean;
                }
                out[b * act_dim + j] = tanhf(v);
            }
    }
};

Agent* make_sac(const bb_sac_cfg& cfg) { return new Sac(cfg); }

}  // namespace bb
