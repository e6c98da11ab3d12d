/*
 * border_host.h -- C entry points of libborder_host.so: the reference's host-side training loops
 * restated in C++ over border_b200.h (border_b200/host/border_host.hpp), runnable without a Rust
 * toolchain.  These are NOT part of the drop-in boundary (that is border_b200.h); they exist so
 * tests and benchmarks can drive the boundary exactly the way border_core::Trainer
 * (border-core/src/trainer.rs:267-327) and border_async_trainer::train_async
 * (border-async-trainer/src/util.rs:31-92) do.
 */
#ifndef BORDER_HOST_H
#define BORDER_HOST_H
#include <stdint.h>
#include "border_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

enum { BBH_ALGO_DQN = 0, BBH_ALGO_IQN = 1, BBH_ALGO_SAC = 2 };

/* Synthetic zero-cost environment. */
typedef struct {
    int32_t obs_kind;        /* BB_U8 | BB_F32 */
    uint32_t obs_elems;
    uint64_t episode_len;    /* is_terminated every episode_len steps (0 = never) */
    uint64_t truncate_len;   /* is_truncated every truncate_len steps (0 = never) */
} bbh_env_cfg;

/* TrainerConfig (border-core/src/trainer/config.rs:30-88) + AsyncTrainerConfig
 * (border-async-trainer/src/async_trainer/config.rs:11-28) + ActorManagerConfig.n_buffer;
 * 0 for an interval means usize::MAX ("never"), as in the reference defaults. */
typedef struct {
    uint64_t max_opts, opt_interval, eval_interval, flush_record_interval, record_compute_cost_interval,
        record_agent_info_interval, warmup_period, save_interval;
    uint64_t sync_interval, n_actors, n_buffer;
    uint64_t env_seed;
} bbh_trainer_cfg;

typedef struct {
    uint64_t env_steps, opt_steps, records, saves, buffer_len, agent_n_opts, samples_total, syncs;
    double opt_seconds, sample_seconds, total_seconds, samples_per_sec, opt_per_sec;
    float last_loss;
} bbh_train_stat;

const char* bbh_last_error(void);
void bbh_trainer_cfg_default(bbh_trainer_cfg* cfg);
/* Trainer::train with a SyntheticEnv, SimpleStepProcessor, a B200 agent (algo + its bb_*_cfg) and a
 * B200 replay buffer. */
int32_t bbh_train(int32_t algo, const void* agent_cfg, const bb_replay_cfg* replay_cfg, const bbh_env_cfg* env_cfg,
                  const bbh_trainer_cfg* trainer_cfg, const char* save_dir, bbh_train_stat* out);
/* Trainer::train_offline (border-core/src/trainer.rs:330-384): max_opts optimisation steps on a replay buffer that already
 * holds the dataset (the caller owns `dataset` and fills it with bb_replay_push); no environment, warmup_period = 0,
 * opt_interval = 1; records / saves at the trainer's intervals. */
int32_t bbh_train_offline(int32_t algo, const void* agent_cfg, bb_replay* dataset, const bbh_trainer_cfg* trainer_cfg,
                          const char* save_dir, bbh_train_stat* out);
/* train_async: n_actors actor threads (each its own agent + env, seed = actor id) -> learner. */
int32_t bbh_train_async(int32_t algo, const void* agent_cfg, const bb_replay_cfg* replay_cfg,
                        const bbh_env_cfg* env_cfg, const bbh_trainer_cfg* trainer_cfg, bbh_train_stat* out);

/* The same with a hook on the learner thread: phase 0 right after the learner agent exists (a data-parallel job -- BASELINE
 * configs[4]: actors per GPU feeding a learner, learners synchronised across GPUs -- connects the learner's gradient peers
 * there, bb_agent_ipc_export / bb_agent_ipc_connect), phase 2 when the replay warm-up is over and the optimisation loop
 * starts, phase 1 after the last optimisation step. */
typedef void (*bbh_learner_hook)(bb_agent* learner, int32_t phase, void* user);
int32_t bbh_train_async_ex(int32_t algo, const void* agent_cfg, const bb_replay_cfg* replay_cfg,
                           const bbh_env_cfg* env_cfg, const bbh_trainer_cfg* trainer_cfg, bbh_learner_hook hook, void* user,
                           bbh_train_stat* out);

/* Test hook (no device work): n_steps x Sampler::sample_and_push (trainer/sampler.rs:99-144 + SimpleStepProcessor::process,
 * step_proc.rs:103-137) with a scripted policy (action of step i = i) and a recording buffer over the synthetic u8 environment;
 * row i of out[n_steps][8] = {obs episode, obs word, next_obs episode, next_obs word, act, reward, is_terminated, is_truncated}
 * of the i-th pushed transition. */
int32_t bbh_sampler_trace(const bbh_env_cfg* env_cfg, uint64_t seed, uint64_t n_steps, int64_t* out);

/* Measurement aid: `n_steps` iterations of the Trainer's inner loop on EXISTING handles, all through
 * the C ABI with host buffers: bb_replay_push(one host transition) + bb_agent_opt(record) -- the
 * loss is read back to the host every step (trainer.rs:206-228 with opt_interval 1).  The n_slots
 * host transitions (obs/next_obs rows of obs_row_bytes, act rows of act_row_bytes) are pushed
 * round-robin.  Returns the last loss. */
int32_t bbh_e2e_steps(bb_agent* agent, bb_replay* replay, const void* obs, const void* act, const void* next_obs,
                      const float* reward, const int8_t* is_terminated, const int8_t* is_truncated,
                      uint64_t obs_row_bytes, uint64_t act_row_bytes, uint64_t n_slots, uint64_t n_steps,
                      float* last_loss);

/* n_steps x Sampler::sample_and_push (border-core/src/trainer/sampler.rs:99-144) with a zero-cost environment:
 * bb_agent_sample on the host observation of slot i % n_slots (discrete action) + bb_replay_push of that transition. */
int32_t bbh_env_steps(bb_agent* agent, bb_replay* replay, const void* obs, const void* next_obs, const float* reward,
                      const int8_t* is_terminated, const int8_t* is_truncated, uint64_t obs_row_bytes, uint64_t n_slots,
                      uint64_t n_steps, int64_t* last_act);
/* bbh_env_steps through bb_actor_step (border_b200.h): one PCIe crossing per step, action chosen on the device. */
int32_t bbh_actor_steps(bb_agent* agent, bb_replay* replay, const void* obs, const void* next_obs, const float* reward,
                      const int8_t* is_terminated, const int8_t* is_truncated, uint64_t obs_row_bytes, uint64_t n_slots,
                      uint64_t n_steps, int64_t* last_act);

#ifdef __cplusplus
}
#endif
#endif
