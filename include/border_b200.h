/*
 * border_b200.h -- C ABI of the B200-native training hot path of laboroai/border.
 *
 * This is the drop-in boundary: what a Rust `border-b200-sys` crate would bind with bindgen so
 * that `impl ReplayBufferBase for B200ReplayBuffer` and `impl Agent<E, R> for B200Dqn/Iqn/Sac`
 * can sit under the unmodified border_core::Trainer and border_async_trainer loops
 * (INTEGRATION.md shows the binding).  Reference citations are relative to /root/reference/.
 *
 * Conventions
 *   - every entry point returns int32_t status, 0 = ok; bb_last_error() gives a thread-local
 *     message.  Nothing unwinds or aborts across the ABI.
 *   - handles are opaque, created/destroyed by the caller, not internally synchronised: one
 *     handle <-> one thread at a time (mirrors `&mut self` in the traits).
 *   - "host" pointers are ordinary (ideally pinned) host memory; "dev" pointers are device memory
 *     on the handle's GPU.  No torch types anywhere.
 *   - all device work of a handle is issued on that handle's CUDA stream
 *     (bb_replay_set_stream / bb_agent_set_stream accept a cudaStream_t as void*).
 */
#ifndef BORDER_B200_H
#define BORDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BB_ABI_VERSION 1

typedef struct bb_replay bb_replay;
typedef struct bb_agent bb_agent;

/* Element kinds of replay rows (tch Kind of the first pushed tensor,
 * border-tch-agent/src/tensor_batch.rs:95-101). */
enum { BB_U8 = 0, BB_F32 = 1, BB_I64 = 2, BB_I32 = 3 };
/* WeightNormalizer, border-core/src/generic_replay_buffer/base/sum_tree.rs:11-18 */
enum { BB_NORM_ALL = 0, BB_NORM_BATCH = 1 };
/* CriticLoss, border-tch-agent/src/util.rs:18-26 */
enum { BB_LOSS_MSE = 0, BB_LOSS_SMOOTH_L1 = 1 };
/* OptimizerConfig, border-tch-agent/src/opt.rs:13-28 */
enum { BB_OPT_ADAM = 0, BB_OPT_ADAMW = 1 };
/* DqnExplorer, border-tch-agent/src/dqn/explorer.rs:9-15 */
enum { BB_EXPLORER_SOFTMAX = 0, BB_EXPLORER_EPS_GREEDY = 1 };
/* network kinds: Mlp (mlp/base.rs), AtariCnn (cnn/base.rs) */
enum { BB_NET_MLP = 0, BB_NET_ATARI_CNN = 1 };
/* EntCoefMode, border-tch-agent/src/sac/ent_coef.rs:13-20 */
enum { BB_ENTCOEF_FIX = 0, BB_ENTCOEF_AUTO = 1 };
/* IqnSample, border-tch-agent/src/iqn/model/base.rs:327-345 */
enum { BB_IQN_CONST10 = 0, BB_IQN_UNIFORM8 = 1, BB_IQN_UNIFORM10 = 2, BB_IQN_UNIFORM32 = 3,
       BB_IQN_UNIFORM64 = 4, BB_IQN_MEDIAN = 5, BB_IQN_CONST32 = 6 };

const char* bb_last_error(void);
int32_t bb_abi_version(void);
int32_t bb_device_count(int32_t* out);
/* Sticky device-side failure flag (0 = none): a bounded wait inside a kernel timed out (tcgen05 / TMA pipelines: 11 producer,
 * 12 MMA issuer, 13 split warps, 14 epilogue; peer barrier: 21).  Every agent entry point checks it and fails loudly. */
int32_t bb_device_error(int32_t* out, int32_t reset);
/* glibc powf restated on the device (sum_tree.rs:76,96,134,139 use f32::powf); test hook. */
int32_t bb_test_powf(int32_t device, const float* host_x, const float* host_y, float* host_out, size_t n);

/* Test hook: one dense GEMM on the device through the SIMT-producer tcgen05 path (use_tc = 1), the fp32
 * CUDA-core path (0), the shape dispatcher the layers use (2: skinny kernels / tcgen05 / CUDA-core tiles) or the
 * TMA-fed tcgen05 kernel (3; operands get their lo planes, and for 2 and 3 the lo plane of C is verified).  mode 0: C = A[M][K] B[N][K]^T (+bias, relu); 2: C = A[M][K] B[K][N];
 * 3: C = A[K][M]^T B[K][N].  All pointers are host memory. */
int32_t bb_test_gemm(int32_t device, int32_t mode, int32_t use_tc, int32_t M, int32_t N, int32_t K,
                     const float* A, const float* B, const float* bias, int32_t relu, float* C_out);

/* Test hook: one convolution layer (NHWC float X[B][H][W][C], weights [OC][k][k][C], stride s, no padding) through the layer
 * primitives of the agents' networks, with (use_tma = 1) or without lo operand planes, i.e. on the TMA-fed im2col tcgen05
 * kernels or the SIMT-producer ones.  mode 0: out = Y[B][OH][OW][OC] (+bias); 1: out = dW[OC][k][k][C] from (dY, X);
 * 2: out = dX[B][H][W][C] from (dY, W).  Host pointers. */
int32_t bb_test_conv(int32_t device, int32_t mode, int32_t use_tma, int32_t B, int32_t C, int32_t H, int32_t W, int32_t OC,
                     int32_t k, int32_t s, const float* X, const float* Wt, const float* bias, const float* dY, float* out);

/* Debug: clock64 stamps of CTA (0,0,0) of the last TMA GEMM launched with BB_TMA_TRACE=1: [3 roles: TMA producer, MMA issuer,
 * split warp][64 k-slices][4 stamps]. */
int32_t bb_tma_trace(int64_t* out);
/* ... and of every CTA (first 1024) of that launch: out[1024][4] = %globaltimer at entry / after the dependency wait / at
 * exit (ns) and the SM id; rows of CTAs that did not exist are zero.  Clears the buffer. */
int32_t bb_tma_trace_ctas(int64_t* out);

/* GEMM launches that took the TMA-fed path / tensor-map constructions the driver refused, since the last reset. */
int32_t bb_tma_stats(uint64_t* launches, uint64_t* rejects, int32_t reset);

/* Timing hook: mean milliseconds of `iters` back-to-back launches of one dense GEMM. */
int32_t bb_bench_gemm(int32_t device, int32_t mode, int32_t use_tc, int32_t M, int32_t N, int32_t K,
                      int32_t iters, float* ms_out);

/* Timing hook: mean milliseconds of `iters` back-to-back launches of the AtariCnn first-layer forward
 * (cnn/base.rs:26-28) on a synthetic u8 batch [B][C][84][84]. */
int32_t bb_bench_conv1(int32_t device, int32_t B, int32_t C, int32_t iters, float* ms_out);

/* Debug hook: clock64 stamps [2 roles][64 k-slices][4 points] of the last traced tcgen05 launch. */
int32_t bb_debug_tc_trace(int64_t* out);

/* ------------------------------------------------------------------------------------------
 * Replay buffer: SimpleReplayBuffer<O, A> (border-core/src/generic_replay_buffer/base.rs:86-426)
 * with TensorBatch storage (border-tch-agent/src/tensor_batch.rs:44-120), as a ring of SoA
 * columns in HBM.
 * ---------------------------------------------------------------------------------------- */

/* SimpleReplayBufferConfig + PerConfig (generic_replay_buffer/config.rs:45-65,185-197); field
 * names are the reference's.  The row geometry replaces TensorBatch's lazy allocation from the
 * first pushed tensor. */
typedef struct {
    uint64_t capacity;
    uint64_t seed;               /* StdRng::seed_from_u64(seed), base.rs:353 */
    int32_t per_config_some;     /* per_config: Option<PerConfig> */
    float alpha, beta_0, beta_final;
    uint64_t n_opts_final;
    int32_t normalize;           /* BB_NORM_* */
    /* row geometry */
    int32_t obs_kind;            /* BB_U8 (Atari frames, lossless) or BB_F32 */
    uint32_t obs_elems;          /* elements per obs row, e.g. 4*84*84 */
    int32_t act_kind;            /* BB_I64 (discrete) or BB_F32 (continuous) */
    uint32_t act_elems;
    /* fastrand is unseeded in the reference (sum_tree.rs:123); here it is a seeded wyrand so
     * that runs are reproducible.  */
    uint64_t fastrand_seed;
    int32_t device;              /* CUDA ordinal */
} bb_replay_cfg;

/* A sampled batch (GenericTransitionBatch, generic_replay_buffer/batch.rs:89-117) as device
 * views into buffers owned by the replay handle; valid until the next call on that handle. */
typedef struct {
    uint64_t batch_size;
    const void* obs;          /* [B, obs_elems] obs_kind */
    const void* act;          /* [B, act_elems] act_kind */
    const void* next_obs;     /* [B, obs_elems] */
    const float* reward;      /* [B] */
    const int8_t* is_terminated; /* [B] */
    const int8_t* is_truncated;  /* [B] */
    const uint64_t* ix_sample;   /* [B]  (Some(ixs), base.rs:398) */
    const float* weight;      /* [B] or NULL when PER is off */
} bb_batch_view;

void    bb_replay_cfg_default(bb_replay_cfg* cfg);                    /* config.rs:199-207 */
int32_t bb_replay_create(const bb_replay_cfg* cfg, bb_replay** out);  /* ReplayBufferBase::build base.rs:336-356 */
int32_t bb_replay_destroy(bb_replay* rb);
int32_t bb_replay_set_stream(bb_replay* rb, void* cuda_stream);
/* ExperienceBufferBase::push (base.rs:295-316) for `n` transitions; host or device sources. */
int32_t bb_replay_push(bb_replay* rb, const void* obs, const void* act, const void* next_obs,
                       const float* reward, const int8_t* is_terminated, const int8_t* is_truncated,
                       size_t n, int32_t src_on_device);
int32_t bb_replay_len(const bb_replay* rb, uint64_t* out);            /* ExperienceBufferBase::len */
/* ReplayBufferBase::batch (base.rs:376-402): fused index generation + gather, all on device. */
int32_t bb_replay_sample(bb_replay* rb, size_t batch_size, bb_batch_view* out);
/* Copies the last sampled batch to host memory (any pointer may be NULL): the `unpack()` view
 * (batch.rs:66-81) for callers that want Vec<usize>/Vec<f32>.  Synchronises the stream. */
int32_t bb_replay_batch_to_host(bb_replay* rb, void* obs, void* act, void* next_obs, float* reward,
                                int8_t* is_terminated, int8_t* is_truncated, uint64_t* ix_sample,
                                float* weight);
/* Rows in the batch the last sample call produced (by bb_replay_sample or inside bb_agent_opt); their indices can then be read
 * with bb_replay_batch_to_host(ix only): `ix_sample` of TransitionBatch::unpack (batch.rs:66-81). */
int32_t bb_replay_last_batch(bb_replay* rb, uint64_t* out);

/* ReplayBufferBase::update_priority (base.rs:413-426).  Pointers are host unless on_device. */
int32_t bb_replay_update_priority(bb_replay* rb, const uint64_t* ixs, const float* td_errs, size_t n,
                                  int32_t on_device);
/* test hooks */
int32_t bb_replay_inject_uniforms(bb_replay* rb, const float* host_u, size_t n); /* replaces fastrand::f32() for the next sample */
int32_t bb_replay_dump_sum_tree(bb_replay* rb, float* host_tree /* 2*capacity-1 */, uint64_t* n_samples, uint64_t* n_opts);
int32_t bb_replay_state(const bb_replay* rb, uint64_t* head_i, uint64_t* size, uint64_t* rng_words_drawn);
/* Fills the whole ring on the device with the synthetic workload of SURVEY.md 8(d)
 * (counter-based generator, seed 1234): benchmark set-up, not part of the path. */
int32_t bb_replay_fill_synthetic(bb_replay* rb, uint64_t n_rows, uint32_t n_actions, uint64_t seed);

/* ------------------------------------------------------------------------------------------
 * Agents: Dqn / Iqn / Sac (border-tch-agent/src/{dqn,iqn,sac}/base.rs) behind
 * Policy + Agent + Configurable + SyncModel (border-core/src/base/{policy,agent}.rs,
 * border-async-trainer/src/sync_model.rs).
 * ---------------------------------------------------------------------------------------- */

/* MlpConfig (mlp/config.rs:7-12) / AtariCnnConfig (cnn/config.rs:13-18) */
typedef struct {
    int32_t kind;             /* BB_NET_* */
    int32_t in_dim;           /* Mlp */
    int32_t n_units;
    int32_t units[8];
    int32_t out_dim;
    int32_t activation_out;
    int32_t n_stack;          /* AtariCnn */
    int32_t skip_linear;
} bb_net_cfg;

/* OptimizerConfig (opt.rs:13-28); Adam{lr} uses tch Adam::default(): beta 0.9/0.999, wd 0, eps 1e-8 */
typedef struct {
    int32_t kind;             /* BB_OPT_* */
    double lr, beta1, beta2, wd, eps;
    int32_t amsgrad;
} bb_opt_cfg;

/* DqnConfig (dqn/config.rs:26-48) + DqnModelConfig (dqn/model/config.rs:12-20) */
typedef struct {
    bb_net_cfg q_config;          /* model_config.q_config */
    bb_opt_cfg opt_config;        /* model_config.opt_config */
    uint64_t soft_update_interval;
    uint64_t n_updates_per_opt;
    uint64_t batch_size;
    double discount_factor;
    double tau;
    int32_t train;
    int32_t explorer;             /* BB_EXPLORER_* */
    double eps_start, eps_final;  /* EpsilonGreedy, dqn/explorer.rs:34-40 */
    uint64_t final_step;
    int32_t clip_reward_some; double clip_reward;   /* stored, unused (dqn/base.rs:42) */
    int32_t double_dqn;
    int32_t clip_td_err_some; double clip_td_err_min, clip_td_err_max;
    int32_t device;               /* Device::Cuda(n) */
    int32_t critic_loss;          /* BB_LOSS_* */
    uint64_t record_verbose_level;
    uint64_t init_seed;           /* weight init generator (libtorch's global RNG in the reference) */
    uint64_t explorer_seed;       /* fastrand seed for exploration */
} bb_dqn_cfg;

/* SacConfig (sac/config.rs:23-47) + ActorConfig / CriticConfig */
typedef struct {
    bb_net_cfg pi_config;         /* actor_config.pi_config (Mlp2: trunk units, out_dim = act dim) */
    bb_opt_cfg pi_opt_config;
    bb_net_cfg q_config;          /* critic_config.q_config (Mlp over [obs, act]) */
    bb_opt_cfg q_opt_config;
    double gamma, tau;
    int32_t ent_coef_mode;        /* BB_ENTCOEF_* */
    double ent_coef_fix;          /* Fix(alpha) */
    double ent_coef_target, ent_coef_lr; /* Auto(target_entropy, lr) */
    double epsilon, min_lstd, max_lstd;
    uint64_t n_updates_per_opt;
    uint64_t batch_size;
    int32_t train;
    int32_t critic_loss;
    double reward_scale;
    uint64_t n_critics;
    int32_t seed_some; int64_t seed;
    int32_t device;
    uint64_t init_seed;
    uint64_t noise_seed;          /* in-kernel Philox for z ~ N(0,1) (CPU randn in the reference) */
} bb_sac_cfg;

/* IqnConfig (iqn/config.rs:21-41) + IqnModelConfig (iqn/model/config.rs:33-51) */
typedef struct {
    bb_net_cfg f_config;          /* feature extractor */
    bb_net_cfg m_config;          /* merge net */
    bb_opt_cfg opt_config;
    int32_t feature_dim, embed_dim;
    uint64_t soft_update_interval, n_updates_per_opt, batch_size;
    double discount_factor, tau;
    int32_t train;
    int32_t sample_percents_pred, sample_percents_tgt, sample_percents_act; /* BB_IQN_* */
    double eps_start, eps_final; uint64_t final_step;   /* IqnExplorer::EpsilonGreedy */
    int32_t device;
    uint64_t init_seed, explorer_seed, tau_seed;
} bb_iqn_cfg;

/* Record of one opt_with_record call (border-core/src/record/base.rs; keys of dqn/base.rs:154,
 * iqn/base.rs:166, sac/base.rs:190-197). */
typedef struct {
    float loss;          /* DQN "loss" */
    float loss_critic;   /* IQN/SAC "loss_critic" */
    float loss_actor;    /* SAC "loss_actor" */
    float ent_coef;      /* SAC "ent_coef" */
    float pred_mean, tgt_mean, reward_mean, tgt_minus_pred_mean;  /* DQN verbose >= 2 */
    uint64_t n_opts;
} bb_record;

void    bb_dqn_cfg_default(bb_dqn_cfg* cfg);  /* dqn/config.rs:82-102 */
void    bb_sac_cfg_default(bb_sac_cfg* cfg);  /* sac/config.rs:85-105 */
void    bb_iqn_cfg_default(bb_iqn_cfg* cfg);  /* iqn/config.rs:50-67 */
int32_t bb_dqn_create(const bb_dqn_cfg* cfg, bb_agent** out);   /* Configurable::build dqn/base.rs:255-287 */
int32_t bb_sac_create(const bb_sac_cfg* cfg, bb_agent** out);   /* sac/base.rs */
int32_t bb_iqn_create(const bb_iqn_cfg* cfg, bb_agent** out);   /* iqn/base.rs */
int32_t bb_agent_destroy(bb_agent* a);
int32_t bb_agent_set_stream(bb_agent* a, void* cuda_stream);
int32_t bb_agent_set_train(bb_agent* a, int32_t train);
/* Tensor-core precision of the agent's large contractions: fast = 0 (default) 3xTF32 with fp32 accumulation -- fp32 parity
 * with the reference's CPU path (1e-4 on losses); fast = 1 one TF32 product per fp32 product (10-bit mantissa operands,
 * fp32 accumulation), ~2-3e-4 relative on the DQN loss (tests/test_fast_mode_gpu.py) for less tensor-core and
 * shared-memory work.  Not part of the reference's surface: an opt-in of this implementation. */
int32_t bb_agent_set_precision(bb_agent* a, int32_t fast);         /* Agent::train / eval */
int32_t bb_agent_is_train(const bb_agent* a, int32_t* out);
/* Policy::sample for `n` observations (host pointers; n = 1 in Sampler::sample_and_push).
 * act_out: int64[n] for DQN/IQN, float[n*act_dim] for SAC. */
int32_t bb_agent_sample(bb_agent* a, const void* obs, size_t n, void* act_out);
/* Sampler::sample_and_push (border-core/src/trainer/sampler.rs:99-144) for one environment step of a discrete-action
 * agent, with the observation crossing PCIe ONCE and the action chosen on the device (dqn/explorer.rs:29-31,68-90; the
 * fastrand draws stay on the host, in the reference's order, so the action sequence is the one bb_agent_sample gives):
 *   `obs`      the observation env.step returned (host), i.e. next_obs of the transition that started at the previous
 *              call's observation with the previous call's action; (reward, is_terminated, is_truncated) belong to it.
 *              That transition is pushed into `rb` from the device-resident copies (nothing is pushed on the first call
 *              after create / bb_actor_reset);
 *   `reset_obs` NULL, or -- when the episode ended -- the observation env.reset() returned: the policy acts on it;
 *   `act_out`  the action for the next env.step (Policy::sample on `reset_obs ? reset_obs : obs`).
 * One host->device copy (the observation row + 6 bytes), one push kernel, the policy forward with the explorer in the
 * tail, an 8-byte action written to pinned host memory, one synchronisation. */
int32_t bb_actor_step(bb_agent* a, bb_replay* rb, const void* obs, const void* reset_obs, float reward, int8_t is_terminated,
                      int8_t is_truncated, int64_t* act_out);
/* The same with `obs` / `reset_obs` already in HBM (device pointers, e.g. bb_atari_obs_device): only reward and flags
 * cross PCIe.  The pointers must stay valid and unchanged until the call returns (it synchronises). */
int32_t bb_actor_step_dev(bb_agent* a, bb_replay* rb, const void* obs_dev, const void* reset_obs_dev, float reward,
                          int8_t is_terminated, int8_t is_truncated, int64_t* act_out);
/* The vectorised form: `n_envs` (1..8) environments per call, as Policy::sample on a batch of n_procs observations does in
 * the reference (dqn/explorer.rs:68-90: one epsilon draw per call, one action draw per process).  obs [n][row], reward /
 * is_terminated / is_truncated [n]; reset_obs [n][row] + reset_mask [n] (rows with mask 0 are ignored) or both NULL;
 * act_out [n].  The n transitions are pushed in environment order.  obs_on_device != 0: obs / reset_obs are device pointers. */
int32_t bb_actor_step_n(bb_agent* a, bb_replay* rb, int32_t n_envs, const void* obs, const void* reset_obs, const int8_t* reset_mask,
                        const float* reward, const int8_t* is_terminated, const int8_t* is_truncated, int64_t* act_out,
                        int32_t obs_on_device);
int32_t bb_actor_reset(bb_agent* a);   /* forget the previous observation (new Sampler / after a manual env.reset) */
/* Agent::opt / opt_with_record (record may be NULL => no device->host copy at all). */
int32_t bb_agent_opt(bb_agent* a, bb_replay* rb, bb_record* record);
int32_t bb_agent_n_opts(const bb_agent* a, uint64_t* out);
/* Measurement aid: one opt() with a CUDA event after every kernel of this library; text_out
 * receives "phase:layer:kernel milliseconds" lines.  bench.py derives the roofline from it. */
int32_t bb_agent_opt_profiled(bb_agent* a, bb_replay* rb, char* text_out, size_t cap);
/* Agent::save_params / load_params (directory of raw named tensors + manifest). */
int32_t bb_agent_save_params(bb_agent* a, const char* dir);
int32_t bb_agent_load_params(bb_agent* a, const char* dir);
/* Named parameter access in the REFERENCE layout (tch VarStore names such as "c1.weight",
 * "mlp.ln0.bias"; conv OIHW, linear [out,in]).  `model` selects the VarStore:
 * DQN: "qnet","qnet_tgt"; IQN: "iqn","iqn_tgt"; SAC: "pi","qnet_0","qnet_tgt_0",...,"ent_coef". */
int32_t bb_agent_param_count(bb_agent* a, const char* model, uint64_t* n_tensors, uint64_t* n_floats);
int32_t bb_agent_param_info(bb_agent* a, const char* model, uint64_t index, char* name_out,
                            size_t name_cap, int64_t* shape_out /* [4] */, int32_t* ndim_out);
int32_t bb_agent_get_param(bb_agent* a, const char* model, const char* name, float* host_out, size_t n);
int32_t bb_agent_set_param(bb_agent* a, const char* model, const char* name, const float* host_in, size_t n);
/* Zeroes the Adam moments and step counters of every model (loading a reference checkpoint, which carries no optimizer
 * state: tch VarStore archives hold the variables only, dqn/base.rs:348-362). */
int32_t bb_agent_reset_opt_state(bb_agent* a);

/* Adam moments of the same tensor (what the reference never saves; for parity tests / resume). */
int32_t bb_agent_get_opt_state(bb_agent* a, const char* model, const char* name, float* host_m,
                               float* host_v, size_t n, uint64_t* step);
/* SyncModel (border-async-trainer/src/sync_model.rs:2-13): model_info() / sync_model() as one flat
 * float blob (DQN: qnet vars; SAC: pi vars, sac/base.rs:377-386). */
int32_t bb_agent_model_info_size(bb_agent* a, uint64_t* n_floats);
int32_t bb_agent_model_info(bb_agent* a, float* host_out, size_t n, uint64_t* n_opts);
int32_t bb_agent_sync_model(bb_agent* a, const float* host_in, size_t n);
/* Same, device to device on one GPU (actor i and learner i share GPU i): pointer-speed sync. */
int32_t bb_agent_sync_model_from(bb_agent* dst, bb_agent* src);
/* Parity hooks: inject the host-RNG draws the reference takes from libtorch's CPU generator.
 * SAC: z ~ N(0,1) [B, act_dim] for update_actor then update_critic (sac/base.rs:76);
 * IQN: tau [B,N] and tau' [B,N'] (iqn/model/base.rs:365-368). */
int32_t bb_agent_inject_noise(bb_agent* a, int32_t slot, const float* host, size_t n);
/* Data-parallel gradient sync (no reference twin).  `peer_grads[r]` is rank r's
 * gradient buffer mapped into this process (CUDA IPC); the fused all-reduce + Adam kernel reads
 * them over NVLink.  world = 1 disables. */
int32_t bb_agent_grad_buffer(bb_agent* a, void** dev_ptr, uint64_t* n_floats);
/* Debug: the gradient exchange's device-side stamps of the last update (32 words: [8..10] FC region start / all ranks
 * ready / done, [12..14] conv region, [16..17] Adam start / every slice delivered; low 32 bits of %globaltimer, ns). */
int32_t bb_agent_exchange_trace(bb_agent* a, uint32_t* out32);
int32_t bb_agent_ipc_export(bb_agent* a, void* handle_out /* 64 bytes grads */, void* flag_handle_out /* 64 bytes */);
int32_t bb_agent_ipc_connect(bb_agent* a, int32_t rank, int32_t world, const void* handles /* world*64 */,
                             const void* flag_handles /* world*64 */);
/* Launch count of this library's kernels since the last reset (bench.py's gpu_launches). */
int32_t bb_kernel_launch_count(uint64_t* out, int32_t reset);

/* ---------------------------------------------------------------------------------------------------------------
 * Atari observation pipeline (border-atari-env/src/env.rs) on the device: RGB24 frames in, the [4][84][84] u8 frame
 * stack (newest first) out, left in HBM for bb_actor_step_dev / bb_replay_push(on_device).
 *   bb_atari_reset   env.rs:263-300  all four frames = warp_and_grayscale(first render)
 *   bb_atari_step    env.rs:126-147 (max of the two last repeated frames), :161-185 (resize 84x84 Triangle + grey),
 *                    :187-199 (stack_frame), :149-159 (clip_reward: sign(r) when built with train != 0)
 * The emulator itself (atari-env-sys) stays on the host and is out of scope. */
typedef struct bb_atari bb_atari;
int32_t bb_atari_create(int32_t device, int32_t width /* 160 */, int32_t height /* 210 */, int32_t train, bb_atari** out);
int32_t bb_atari_destroy(bb_atari* st);
int32_t bb_atari_set_stream(bb_atari* st, void* cuda_stream);
int32_t bb_atari_reset(bb_atari* st, const uint8_t* rgb /* host [h][w][3] */);
int32_t bb_atari_step(bb_atari* st, const uint8_t* rgb_a, const uint8_t* rgb_b, float reward, float* reward_out);
/* current observation: device pointer (valid until the next-but-one step; `consumer_stream` is made to wait for it) ... */
int32_t bb_atari_obs_device(bb_atari* st, void* consumer_stream, const uint8_t** dev_ptr);
/* ... or a host copy of the 28,224 bytes (synchronises) */
int32_t bb_atari_obs_host(bb_atari* st, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* BORDER_B200_H */
