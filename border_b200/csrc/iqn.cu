// iqn.cu -- Iqn agent (border-tch-agent/src/iqn/base.rs) on the device.
//
//   update_critic        iqn/base.rs:63-170        opt_         iqn/base.rs:172-190
//   IqnModel::forward    iqn/model/base.rs:198-234 cos_embed_nn iqn/model/base.rs:162-191
//   IqnSample            iqn/model/base.rs:327-387 average      iqn/model/base.rs:394-418
//   quantile_huber_loss  util/quantile_loss.rs:7-12
//   Policy::sample       iqn/base.rs:204-228, IqnExplorer::EpsilonGreedy iqn/explorer.rs:78-97
//
//   z(x, tau) = f( psi(x)[:, None, :] * relu(W cos(pi i tau) + b) ),  i = 1..embed_dim
//   loss = mean_{b,n',n} |tau[b,n] - 1{d<0}| huber(d),  d = tgt[b,n'] - pred[b,n]
// IQN ignores PER outputs (`_ixs, _weight`, iqn/base.rs:66; update_priority is commented out).
#include <math.h>
#include "agent.cuh"

namespace bb {

__device__ __forceinline__ unsigned long long iqn_mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// IqnSample::sample (iqn/model/base.rs:347-372) -> tau [B][N]; uniform modes draw U[0,1) in the
// kernel (the reference uses Tensor::rand on the CPU generator; parity tests inject tau).
// ctr_dev != null: the counter is read from device memory (the update's launches are argument-invariant: CUDA graph)
__global__ void iqn_tau_kernel(float* tau, int B, int N, int mode, unsigned long long seed, unsigned long long ctr0,
                               const unsigned long long* __restrict__ ctr_dev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    if (ctr_dev) ctr0 = *ctr_dev;
    int n = i % N;
    float t;
    switch (mode) {
        case BB_IQN_CONST10: t = 0.05f + 0.1f * (float)n; break;
        case BB_IQN_CONST32: t = (1.0f / 32.0f) * (float)n; break;  // (1/32) * range(0, 32): 33 points
        case BB_IQN_MEDIAN: t = 0.5f; break;
        default: t = (float)(uint32_t)(iqn_mix64(seed ^ ((ctr0 + i) * 0xD1342543DE82EF95ull)) >> 40) * (1.0f / 16777216.0f);
    }
    tau[i] = t;
}

// cos(tau * (pi * i)), i = 1..E  -> [B*N][E]   (iqn/model/base.rs:170-179)
__global__ void iqn_set_ctr_kernel(unsigned long long a, unsigned long long b, unsigned long long* dst) {
    if (threadIdx.x == 0) { dst[0] = a; dst[1] = b; }
}

__global__ void iqn_cos_kernel(const float* __restrict__ tau, float* __restrict__ out, int BN, int E) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BN * E) return;
    int r = i / E, e = i % E;
    float pi_i = __fmul_rn(3.14159265358979323846f, (float)(e + 1));
    out[i] = cosf(__fmul_rn(tau[r], pi_i));
}

// m[b,n,:] = psi[b,:] * phi[b,n,:]   (iqn/model/base.rs:218-224)
__global__ void iqn_merge_kernel(const float* __restrict__ psi, const float* __restrict__ phi, float* __restrict__ m,
                                 int B, int N, int F) {
    size_t total = (size_t)B * N * F / 4;
    const int F4 = F / 4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int f4 = (int)(i % F4);
        int b = (int)(i / ((size_t)N * F4));
        float4 p = __ldg(reinterpret_cast<const float4*>(psi + (size_t)b * F) + f4);
        float4 q = __ldg(reinterpret_cast<const float4*>(phi) + i);
        reinterpret_cast<float4*>(m)[i] = make_float4(p.x * q.x, p.y * q.y, p.z * q.z, p.w * q.w);
    }
}

// backward of the merge: dphi = dm * psi ; dpsi[b,f] = sum_n dm[b,n,f] * phi[b,n,f]
__global__ void iqn_merge_bwd_kernel(const float* __restrict__ dm, const float* __restrict__ psi,
                                     const float* __restrict__ phi, float* __restrict__ dphi, float* __restrict__ dpsi,
                                     int B, int N, int F) {
    size_t total = (size_t)B * F;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(i / F), f = (int)(i % F);
        float p = psi[i], acc = 0.f;
        for (int n = 0; n < N; ++n) {
            size_t j = ((size_t)b * N + n) * F + f;
            float g = dm[j];
            acc += g * phi[j];
            dphi[j] = g * p;
        }
        dpsi[i] = acc;
    }
}

// One CTA per batch row: target (iqn/base.rs:117-152), quantile Huber loss and its gradient
// (iqn/base.rs:154-164, quantile_loss.rs:7-12).
__global__ void __launch_bounds__(256) iqn_loss_kernel(const float* __restrict__ z, const float* __restrict__ zt,
                                                       const float* __restrict__ tau, const long long* __restrict__ act,
                                                       int act_stride, const float* __restrict__ reward,
                                                       const int8_t* __restrict__ term, float* __restrict__ dz,
                                                       float* __restrict__ loss_part, int B, int N, int Nt, int A,
                                                       float gamma) {
    extern __shared__ float sm[];
    float* s_pred = sm;            // [N]
    float* s_tau = s_pred + N;     // [N]
    float* s_tgt = s_tau + N;      // [Nt]
    float* s_mean = s_tgt + Nt;    // [A]
    float* s_red = s_mean + A;     // [blockDim]
    __shared__ int s_best;
    const int b = blockIdx.x, t = threadIdx.x;
    const int a = (int)act[(size_t)b * act_stride];
    for (int n = t; n < N; n += blockDim.x) {
        s_pred[n] = z[((size_t)b * N + n) * A + a];  // z.gather(-1, a)
        s_tau[n] = tau[(size_t)b * N + n];
    }
    for (int j = t; j < A; j += blockDim.x) {  // y = z'.mean(1)
        float acc = 0.f;
        for (int n = 0; n < Nt; ++n) acc += zt[((size_t)b * Nt + n) * A + j];
        s_mean[j] = acc / (float)Nt;
    }
    __syncthreads();
    if (t == 0) {
        int best = 0;
        for (int j = 1; j < A; ++j)
            if (s_mean[j] > s_mean[best]) best = j;
        s_best = best;
    }
    __syncthreads();
    const float nt = (float)(1 - (int)term[b]);
    const float r = reward[b];
    for (int n = t; n < Nt; n += blockDim.x)  // reward + (1 - is_terminated) * discount_factor * z
        s_tgt[n] = __fadd_rn(r, __fmul_rn(__fmul_rn(nt, gamma), zt[((size_t)b * Nt + n) * A + s_best]));
    __syncthreads();
    const float scale = 1.0f / ((float)B * (float)Nt * (float)N);
    float lsum = 0.f;
    for (int n = t; n < N; n += blockDim.x) {
        const float pred = s_pred[n], tq = s_tau[n];
        float g = 0.f;
        for (int k = 0; k < Nt; ++k) {
            float d = s_tgt[k] - pred;
            float ad = fabsf(d);
            float w = fabsf(tq - (d < 0.f ? 1.f : 0.f));
            lsum += w * (ad < 1.f ? 0.5f * d * d : ad - 0.5f);
            g += w * (ad < 1.f ? d : (d > 0.f ? 1.f : -1.f));
        }
        // d(diff)/d(pred) = -1
        float* dzr = dz + ((size_t)b * N + n) * A;
        for (int j = 0; j < A; ++j) dzr[j] = (j == a) ? -g * scale : 0.f;
    }
    s_red[t] = lsum;
    __syncthreads();
    for (int k = blockDim.x / 2; k > 0; k >>= 1) {
        if (t < k) s_red[t] += s_red[t + k];
        __syncthreads();
    }
    if (t == 0) loss_part[b] = s_red[0] * scale;
}

__global__ void __launch_bounds__(1024) iqn_loss_final_kernel(const float* part, int B, float* out) {
    __shared__ float s[1024];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) acc += part[b];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 512; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0];
}

// mean over the percent points: [B][N][A] -> [B][A]   (average(), iqn/model/base.rs:405-408)
__global__ void iqn_mean_kernel(const float* __restrict__ z, float* __restrict__ out, int B, int N, int A) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * A) return;
    int b = i / A, j = i % A;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc += z[((size_t)b * N + n) * A + j];
    out[i] = acc / (float)N;
}

static int n_percent_points(int mode, bool act_mode) {
    switch (mode) {
        case BB_IQN_CONST10: return 10;
        case BB_IQN_UNIFORM8: return 8;
        case BB_IQN_UNIFORM10: return 10;
        case BB_IQN_UNIFORM32: return 32;
        case BB_IQN_UNIFORM64: return 64;
        case BB_IQN_MEDIAN: return 1;
        case BB_IQN_CONST32:
            // Const32 builds 33 points but reports 32 (iqn/model/base.rs:353-356,376): the reference's
            // update_critic panics on the shape mismatch, only `average` (mean over dim 1) survives it.
            if (!act_mode) throw Error("IqnSample::Const32 is unusable for pred/tgt (33 points vs n_percent_points 32)");
            return 33;
    }
    throw Error("unknown IqnSample");
}

struct IqnWs {
    NetWorkspace f, phi, m;
    float *tau = nullptr, *cos = nullptr, *merged = nullptr, *dmerged = nullptr;
    int B = 0, N = 0;
    void release() {
        f.release(); phi.release(); m.release();
        cudaFree(tau); cudaFree(cos); cudaFree(merged); cudaFree(dmerged);
        tau = cos = merged = dmerged = nullptr;
    }
};

struct Iqn : Agent {
    bb_iqn_cfg cfg;
    Net f_net, phi_net, m_net;
    size_t base_f = 0, base_phi = 0, base_m = 0;
    Model iqn, iqn_tgt;
    IqnWs ws_online, ws_tgt, ws_act;
    int F, E, A;
    float *d_loss_part = nullptr, *d_out = nullptr, *d_qmean = nullptr, *d_qmean_actor = nullptr;
    float* d_inject[2] = {nullptr, nullptr};
    size_t inject_n[2] = {0, 0};
    uint64_t tau_ctr = 0, soft_update_counter = 0, eps_n_opts = 0;
    unsigned long long* d_ctr = nullptr;   // tau counters of the update's two forwards (device memory)
    UpdateGraph graph;
    FastRand fr;
    uint8_t *d_obs_in = nullptr, *h_obs_in = nullptr;
    float* h_q = nullptr;
    size_t obs_in_cap = 0;

    explicit Iqn(const bb_iqn_cfg& c) : cfg(c), fr(c.explorer_seed) {
        init_base(c.device);
        DeviceGuard g(device);
        train = c.train != 0;
        F = c.feature_dim; E = c.embed_dim;
        BB_CHECK(F >= 4 && F % 4 == 0 && E >= 1, "feature_dim must be a multiple of 4");
        f_net.build(c.f_config, "");
        BB_CHECK(f_net.out_dim == F, "feature extractor output must equal feature_dim");
        f_net.init_tables(device);
        phi_net.reset();
        phi_net.add_linear_layer("iqn_cos_to_feature", E, F, true);
        phi_net.in_elems = E;
        BB_CHECK(c.m_config.kind == BB_NET_MLP && c.m_config.in_dim == F, "merge net must be an Mlp over feature_dim");
        m_net.build(c.m_config, "");
        A = m_net.out_dim;
        // feature order: AtariCnn.skip_linear flattens (c,h,w) in the reference, (h,w,c) here; the
        // tensors that meet that axis are permuted on import/export.
        if (c.f_config.kind == BB_NET_ATARI_CNN) {
            phi_net.params[0].perm = 3; phi_net.params[0].pc = 64; phi_net.params[0].ph = 7; phi_net.params[0].pw = 7;
            phi_net.params[1].perm = 4; phi_net.params[1].pc = 64; phi_net.params[1].ph = 7; phi_net.params[1].pw = 7;
            m_net.params[0].perm = 2; m_net.params[0].pc = 64; m_net.params[0].ph = 7; m_net.params[0].pw = 7;
        }
        // one VarStore: psi vars, then iqn_cos_to_feature, then the merge net (iqn/model/base.rs:64-86)
        base_f = 0; base_phi = f_net.n_params; base_m = base_phi + phi_net.n_params;
        for (Model* m : {&iqn, &iqn_tgt}) {
            m->params.clear();
            for (auto pi : f_net.params) { pi.offset += base_f; m->params.push_back(pi); }
            for (auto pi : phi_net.params) { pi.offset += base_phi; m->params.push_back(pi); }
            for (auto pi : m_net.params) { pi.offset += base_m; m->params.push_back(pi); }
            m->n = base_m + m_net.n_params;
        }
        iqn.name = "iqn"; iqn.alloc(true); iqn.set_hyper(c.opt_config);
        iqn_tgt.name = "iqn_tgt"; iqn_tgt.alloc(false);
        f_net.init_params(ctx, iqn.p + base_f, c.init_seed * 31 + 1);
        phi_net.init_params(ctx, iqn.p + base_phi, c.init_seed * 31 + 2);
        m_net.init_params(ctx, iqn.p + base_m, c.init_seed * 31 + 3);
        iqn_tgt.copy_params_from(iqn, ctx.stream);
        models = {&iqn, &iqn_tgt};
        d_out = dev_alloc_zero<float>(8, ctx.stream);
        d_ctr = dev_alloc_zero<unsigned long long>(2, ctx.stream);
        BB_CUDA(cudaStreamSynchronize(ctx.stream));
    }
    ~Iqn() override {
        DeviceGuard g(device);
        cudaStreamSynchronize(ctx.stream);
        graph.reset();
        cudaFree(d_ctr);
        ws_online.release(); ws_tgt.release(); ws_act.release();
        iqn.release(); iqn_tgt.release();
        f_net.free_tables();
        cudaFree(d_loss_part); cudaFree(d_out); cudaFree(d_qmean); cudaFree(d_qmean_actor); cudaFree(d_inject[0]); cudaFree(d_inject[1]);
        cudaFree(d_obs_in);
        if (h_obs_in) cudaFreeHost(h_obs_in);
        if (h_q) cudaFreeHost(h_q);
    }
    Model* sync_model_src() override { return &iqn; }
    void precision_changed() override { graph.reset(); }
    void grad_buffer(void** p, uint64_t* n) override { *p = iqn.g; *n = iqn.n; }

    void inject_noise(int slot, const float* host, size_t n) override {
        DeviceGuard g(device);
        BB_CHECK(slot == 0 || slot == 1, "IQN noise slots: 0 = tau (pred), 1 = tau' (target)");
        BB_CUDA(cudaStreamSynchronize(ctx.stream));
        cudaFree(d_inject[slot]);
        d_inject[slot] = dev_alloc<float>(n);
        h2d_sync(d_inject[slot], host, n * 4, ctx.stream);
        inject_n[slot] = n;
    }

    void ensure(IqnWs& w, int B, int N, bool with_grad) {
        if (B <= w.B && N <= w.N) return;
        BB_CUDA(cudaStreamSynchronize(ctx.stream));
        w.release();
        f_net.alloc_workspace(w.f, B, with_grad);
        phi_net.alloc_workspace(w.phi, B * N, with_grad);
        m_net.alloc_workspace(w.m, B * N, with_grad);
        w.tau = dev_alloc<float>((size_t)B * N);
        w.cos = dev_alloc<float>((size_t)B * N * E);
        w.merged = dev_alloc<float>((size_t)B * N * F);
        if (with_grad) w.dmerged = dev_alloc<float>((size_t)B * N * F);
        w.B = B; w.N = N;
    }

    // IqnModel::forward (iqn/model/base.rs:198-234) -> z [B*N][A]
    const float* forward(const Model& mdl, const void* obs, int B, int N, int mode, int slot, IqnWs& w) {
        if (slot >= 0 && inject_n[slot]) {
            BB_CHECK(inject_n[slot] == (size_t)B * N, "injected tau has the wrong size");
            BB_CUDA(cudaMemcpyAsync(w.tau, d_inject[slot], (size_t)B * N * 4, cudaMemcpyDeviceToDevice, ctx.stream));
        } else {
            // (slots 0 / 1 = the update's two forwards: their counters were written to d_ctr by update_critic)
            iqn_tau_kernel<<<(B * N + 255) / 256, 256, 0, ctx.stream>>>(w.tau, B, N, mode, cfg.tau_seed, tau_ctr,
                                                                        slot >= 0 ? d_ctr + slot : nullptr);
            BB_LAUNCHED();
            if (slot < 0) tau_ctr += (uint64_t)B * N;
        }
        const float* psi = f_net.forward(ctx, mdl.p + base_f, obs, f_net.in_elems, B, w.f);
        iqn_cos_kernel<<<(B * N * E + 255) / 256, 256, 0, ctx.stream>>>(w.tau, w.cos, B * N, E);
        BB_LAUNCHED();
        ctx.layer = "cos"; ctx.mark("iqn_cos");
        const float* phi = phi_net.forward(ctx, mdl.p + base_phi, w.cos, E, B * N, w.phi);
        size_t tot4 = (size_t)B * N * F / 4;
        iqn_merge_kernel<<<(int)std::min<size_t>((tot4 + 255) / 256, (size_t)ctx.sms * 16), 256, 0, ctx.stream>>>(psi, phi, w.merged, B, N, F);
        BB_LAUNCHED();
        ctx.layer = "merge"; ctx.mark("iqn_merge");
        return m_net.forward(ctx, mdl.p + base_m, w.merged, F, B * N, w.m);
    }

    float update_critic(Replay& rb, bool want_loss) {
        const int B = (int)cfg.batch_size;
        const int N = n_percent_points(cfg.sample_percents_pred, false);
        const int Nt = n_percent_points(cfg.sample_percents_tgt, false);
        BB_CHECK(B >= 1 && B <= 65535, "batch_size out of range");
        ensure(ws_online, B, N, true);
        ensure(ws_tgt, B, Nt, false);
        if (!d_loss_part) d_loss_part = dev_alloc<float>(65536);
        BB_CHECK(rb.obs_row_bytes == (uint32_t)f_net.in_elems * (f_net.u8_input ? 1u : 4u),
                 "replay obs rows do not match the feature extractor input");
        BB_CHECK(rb.cfg.act_kind == BB_I64, "IQN needs i64 action rows");
        // this update's tau counters -> device memory (one launch), so that everything below is argument-invariant
        const uint64_t c0 = tau_ctr;
        if (!inject_n[0]) tau_ctr += (uint64_t)B * N;
        const uint64_t c1 = tau_ctr;
        if (!inject_n[1]) tau_ctr += (uint64_t)B * Nt;
        iqn_set_ctr_kernel<<<1, 32, 0, ctx.stream>>>(c0, c1, d_ctr);
        BB_LAUNCHED();
        bb_batch_view bv;
        graph.run(ctx, rb, B, inject_n[0] == 0 && inject_n[1] == 0, [&]() { enqueue_update(rb, B, N, Nt, bv, true); },
                  [&]() { rb.sample(B, &bv, false); });
        iqn.step += 1;
        ctx.phase = "optimizer";
        synced_adam(iqn);
        inject_n[0] = inject_n[1] = 0;
        if (want_loss) {
            BB_CUDA(cudaMemcpyAsync(h_scratch, d_out, 4, cudaMemcpyDeviceToHost, ctx.stream));
            BB_CUDA(cudaStreamSynchronize(ctx.stream));
            return h_scratch[0];
        }
        return 0.f;
    }

    // every launch of one update up to the optimizer (update_critic, iqn/base.rs:110-170); argument-invariant
    void enqueue_update(Replay& rb, int B, int N, int Nt, bb_batch_view& bv, bool launch_sample) {
        if (rb.stream != ctx.stream) stream_wait(rb.stream, ctx.stream);
        rb.sample(B, &bv, launch_sample);
        if (rb.stream != ctx.stream) stream_wait(ctx.stream, rb.stream);
        ctx.phase = "replay"; ctx.layer = "batch"; ctx.mark("sample_gather");
        ctx.phase = "fwd_online";
        const float* z = forward(iqn, bv.obs, B, N, cfg.sample_percents_pred, 0, ws_online);
        ctx.phase = "fwd_target";
        const float* zt = forward(iqn_tgt, bv.next_obs, B, Nt, cfg.sample_percents_tgt, 1, ws_tgt);
        size_t smem = (size_t)(2 * N + Nt + A + 256) * sizeof(float);
        iqn_loss_kernel<<<B, 256, smem, ctx.stream>>>(z, zt, ws_online.tau, (const long long*)bv.act, (int)rb.cfg.act_elems,
                                                      bv.reward, bv.is_terminated, ws_online.m.dact.back(), d_loss_part, B,
                                                      N, Nt, A, (float)cfg.discount_factor);
        BB_LAUNCHED();
        iqn_loss_final_kernel<<<1, 1024, 0, ctx.stream>>>(d_loss_part, B, d_out);
        BB_LAUNCHED();
        ctx.phase = "loss"; ctx.layer = "quantile_huber"; ctx.mark("iqn_loss");
        // backward: merge net -> (phi, psi) -> cos-embedding linear, feature extractor
        ctx.phase = "backward";
        m_net.backward(ctx, iqn.p + base_m, iqn.g + base_m, ws_online.merged, F, B * N, ws_online.m, ws_online.dmerged, F);
        size_t tot = (size_t)B * F;
        iqn_merge_bwd_kernel<<<(int)std::min<size_t>((tot + 127) / 128, (size_t)ctx.sms * 16), 128, 0, ctx.stream>>>(
            ws_online.dmerged, ws_online.f.act.back(), ws_online.phi.act.back(), ws_online.phi.dact.back(),
            ws_online.f.dact.back(), B, N, F);
        BB_LAUNCHED();
        ctx.layer = "merge"; ctx.mark("iqn_merge_bwd");
        phi_net.backward(ctx, iqn.p + base_phi, iqn.g + base_phi, ws_online.cos, E, B * N, ws_online.phi, nullptr, 0);
        f_net.backward(ctx, iqn.p + base_f, iqn.g + base_f, bv.obs, f_net.in_elems, B, ws_online.f, nullptr, 0);
    }

    void opt(Replay& rb, bb_record* rec) override {  // opt_, iqn/base.rs:172-190
        DeviceGuard g(device);
        float loss = 0.f;
        for (uint64_t i = 0; i < cfg.n_updates_per_opt; ++i) loss += update_critic(rb, rec != nullptr);
        soft_update_counter += 1;
        if (soft_update_counter == cfg.soft_update_interval) {
            soft_update_counter = 0;
            ctx.phase = "target_update";
            track(ctx, iqn_tgt.p, iqn.p, iqn.n, cfg.tau);
        }
        n_opts += 1;
        if (rec) {
            memset(rec, 0, sizeof(*rec));
            rec->loss_critic = loss / (float)cfg.n_updates_per_opt;
            rec->n_opts = n_opts;
        }
    }

    // Policy::sample (iqn/base.rs:204-228)
    // ---- device-side actor path (Agent::actor_step): quantile mean on the device, explorer draws on the host
    size_t actor_obs_row_bytes() const override { return (size_t)f_net.in_elems * (f_net.u8_input ? 1 : 4); }
    int actor_n_actions() const override { return A; }
    const float* actor_q(const uint8_t* d_obs, int n) override {
        const int N = n_percent_points(cfg.sample_percents_act, true);
        ensure(ws_act, std::max(n, ws_act.B), N, false);
        if (!d_qmean_actor) d_qmean_actor = dev_alloc<float>((size_t)kActorMaxEnvs * A);
        const float* z = forward(iqn, d_obs, n, N, cfg.sample_percents_act, -1, ws_act);
        iqn_mean_kernel<<<(n * A + 127) / 128, 128, 0, ctx.stream>>>(z, d_qmean_actor, n, N, A);
        BB_LAUNCHED();
        return d_qmean_actor;
    }
    void actor_pick(int n, ActorPick* out) override {   // IqnExplorer::EpsilonGreedy::action (iqn/explorer.rs:78-97); eval: argmax
        for (int i = 0; i < n; ++i) out[i] = ActorPick{};
        if (train) {
            double d = (cfg.eps_start - cfg.eps_final) / (double)cfg.final_step;
            double eps = std::max(cfg.eps_start - d * (double)eps_n_opts, cfg.eps_final);
            const bool is_random = fr.f64() < eps;
            eps_n_opts += 1;
            if (is_random)
                for (int i = 0; i < n; ++i) { out[i].mode = 1; out[i].forced = (long long)fr.u32_below((uint32_t)A); }
        }
    }

    void sample(const void* obs, size_t n, void* act_out) override {
        DeviceGuard g(device);
        BB_CHECK(n >= 1 && n <= 1024, "sample: n out of range");
        const int N = n_percent_points(cfg.sample_percents_act, true);
        size_t row = (size_t)f_net.in_elems * (f_net.u8_input ? 1 : 4);
        ensure(ws_act, (int)n, N, false);
        if (n * row > obs_in_cap) {
            BB_CUDA(cudaStreamSynchronize(ctx.stream));
            cudaFree(d_obs_in); cudaFree(d_qmean);
            if (h_obs_in) cudaFreeHost(h_obs_in);
            if (h_q) cudaFreeHost(h_q);
            obs_in_cap = n * row;
            d_obs_in = dev_alloc<uint8_t>(obs_in_cap);
            d_qmean = dev_alloc<float>(n * A);
            BB_CUDA(cudaMallocHost(&h_obs_in, obs_in_cap));
            BB_CUDA(cudaMallocHost(&h_q, n * A * sizeof(float)));
        }
        memcpy(h_obs_in, obs, n * row);
        BB_CUDA(cudaMemcpyAsync(d_obs_in, h_obs_in, n * row, cudaMemcpyHostToDevice, ctx.stream));
        const float* z = forward(iqn, d_obs_in, (int)n, N, cfg.sample_percents_act, -1, ws_act);
        iqn_mean_kernel<<<((int)n * A + 127) / 128, 128, 0, ctx.stream>>>(z, d_qmean, (int)n, N, A);
        BB_LAUNCHED();
        BB_CUDA(cudaMemcpyAsync(h_q, d_qmean, n * A * sizeof(float), cudaMemcpyDeviceToHost, ctx.stream));
        BB_CUDA(cudaStreamSynchronize(ctx.stream));
        int64_t* out = (int64_t*)act_out;
        auto argmax = [&](size_t i) {
            int best = 0;
            for (int j = 1; j < A; ++j)
                if (h_q[i * A + j] > h_q[i * A + best]) best = j;
            return (int64_t)best;
        };
        if (train) {  // IqnExplorer::EpsilonGreedy::action (iqn/explorer.rs:78-97)
            double d = (cfg.eps_start - cfg.eps_final) / (double)cfg.final_step;
            double eps = std::max(cfg.eps_start - d * (double)eps_n_opts, cfg.eps_final);
            bool is_random = fr.f64() < eps;
            eps_n_opts += 1;
            for (size_t i = 0; i < n; ++i) out[i] = is_random ? (int64_t)fr.u32_below((uint32_t)A) : argmax(i);
        } else {
            for (size_t i = 0; i < n; ++i) out[i] = argmax(i);
        }
    }
};

Agent* make_iqn(const bb_iqn_cfg& cfg) { return new Iqn(cfg); }

}  // namespace bb
