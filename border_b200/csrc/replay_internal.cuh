// replay_internal.cuh -- the Replay object behind the opaque bb_replay handle (shared with the
// agents, which read the sampled batch straight from its device buffers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/border_b200.h"

namespace bb {

struct ChaChaKey { uint32_t k[8]; };

// Device-resident counters of SimpleReplayBuffer / SumTree / IwScheduler / the RNGs.  Kernels read
// and advance them; the host mirrors them (they evolve deterministically).
struct ReplayCtl {
    unsigned long long rng_pos;    // StdRng words drawn (base.rs:386)
    unsigned long long size;       // SimpleReplayBuffer.size
    unsigned long long head;       // SimpleReplayBuffer.i
    unsigned long long n_samples;  // SumTree.n_samples
    unsigned long long n_opts;     // IwScheduler.n_opts
    unsigned long long fr_draws;   // fastrand draws consumed
    unsigned int done;             // last-CTA-done counter
    unsigned int inject_n;         // injected uniforms pending (test hook)
    float push_p;                  // sum_tree.max() captured by set_priority
    float pad;
};

struct PerParams;
constexpr size_t kMaxInject = 65536;
constexpr int kStageSlots = 4;

struct Replay {
    bb_replay_cfg cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t obs_row_bytes = 0, act_row_bytes = 0, chunk_bytes = 0, n_chunks = 1;
    int vec = 1;
    bool per = false;
    int powf_fused = 1;
    ChaChaKey key;
    // ring columns
    uint8_t *obs = nullptr, *next_obs = nullptr, *act = nullptr;
    float* reward = nullptr;
    int8_t *term = nullptr, *trunc = nullptr;
    // PER
    float *tree = nullptr, *min_tree = nullptr, *max_tree = nullptr, *inject_u = nullptr;
    ReplayCtl* ctl = nullptr;
    // host mirror of ReplayCtl
    uint64_t head = 0, size = 0, n_samples = 0, n_opts = 0, rng_pos = 0, fr_draws = 0, inject_pending = 0;
    // last sampled batch
    size_t batch_cap = 0, last_batch = 0;
    uint64_t batch_generation = 0;   // bumped whenever the batch buffers are reallocated: captured CUDA graphs key on it
    uint8_t *b_obs = nullptr, *b_next_obs = nullptr, *b_act = nullptr;
    float* b_reward = nullptr;
    int8_t *b_term = nullptr, *b_trunc = nullptr;
    unsigned long long* b_ix = nullptr;
    float* b_weight = nullptr;
    // push staging (pinned host ring + device ring)
    uint8_t *stage_host = nullptr, *stage_dev = nullptr;
    size_t stage_cap = 0;
    int stage_slot = 0;
    cudaEvent_t stage_events[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    // host pushes: the H2D copy of the staging block runs on its own stream (copy engine) and the push kernel on
    // `stream` waits for its event, so the copy overlaps whatever the stream is still computing (the previous update)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_events[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool stage_busy[kStageSlots] = {false, false, false, false};
    // update_priority staging
    unsigned long long* upd_ix = nullptr;
    float* upd_td = nullptr;
    size_t upd_cap = 0;

    explicit Replay(const bb_replay_cfg& c);
    ~Replay();
    Replay(const Replay&) = delete;
    Replay& operator=(const Replay&) = delete;

    void ensure_batch(size_t B);
    void push(const void* o, const void* a, const void* no, const float* r, const int8_t* t, const int8_t* tr,
              size_t n, bool on_device);
    // launch = false only advances the host mirror and fills `out` (the launch itself is replayed by a
    // CUDA graph that captured an identical call: every kernel argument of sample() is launch-invariant)
    // gather_obs = false: indices + small columns only; out->obs / next_obs are then the RING columns, to be read through
    // out->ix_sample (the DQN update's first-layer kernels do: no 14.5 MB batch is written and re-read)
    void sample(size_t B, bb_batch_view* out, bool launch = true, bool gather_obs = true);
    void update_priority_dev(const unsigned long long* ixs, const float* td, size_t n);
    void update_priority_host(const uint64_t* ixs, const float* td, size_t n);
    void fill_synthetic(uint64_t n_rows, uint32_t n_actions, uint64_t seed);
    PerParams per_params() const;
    void launch_per_update(const unsigned long long* ixs, const float* td, size_t n, int mode, bool bump);
};

}  // namespace bb

// the opaque handle of include/border_b200.h
struct bb_replay {
    bb::Replay impl;
    explicit bb_replay(const bb_replay_cfg& c) : impl(c) {}
};
