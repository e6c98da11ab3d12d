// host_capi.cpp -- C entry points of libborder_host.so: border's sync / async training loops
// (border_host.hpp) over the C ABI of libborder_b200.so, with synthetic environments.  Built with
// plain g++ (no CUDA headers): everything device-side goes through include/border_b200.h.
#include <string.h>
#include "border_host.hpp"
#include "border_host.h"

using namespace border;

namespace {

thread_local std::string g_err;

uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Zero-cost synthetic environment: observations are a pure function of
// (seed, episode, t); an episode terminates every `episode_len` steps; reward in {-1, 0, 1}.
class SyntheticEnv : public Env {
  public:
    SyntheticEnv(const bbh_env_cfg& c, uint64_t seed) : c_(c), seed_(seed) {}
    Bytes reset() override {
        episode_ += 1;
        t_ = 0;
        return make_obs();
    }
    Step step(const Bytes& act) override {
        // BBH_ENV_STEP_US: emulated cost of one environment step (default 0 = free).  bench.py --topology async sets it to an
        // ALE-like figure so that the actors of BASELINE configs[4] produce at a realistic rate instead of flooding the learner.
        static const long step_us = getenv("BBH_ENV_STEP_US") ? atol(getenv("BBH_ENV_STEP_US")) : 0;
        if (step_us > 0) {
            const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(step_us);
            while (std::chrono::steady_clock::now() < until) {}
        }
        t_ += 1;
        steps_ += 1;
        Step s;
        s.act = act;
        s.obs = make_obs();
        uint64_t h = mix64(seed_ ^ (steps_ * 0x100000001B3ull));
        uint64_t r = h % 100;
        s.reward = r < 5 ? -1.f : (r < 95 ? 0.f : 1.f);
        s.is_terminated = (c_.episode_len && t_ % c_.episode_len == 0) ? 1 : 0;
        s.is_truncated = (!s.is_terminated && c_.truncate_len && t_ % c_.truncate_len == 0) ? 1 : 0;
        return s;
    }
  private:
    Bytes make_obs() {
        size_t bytes = (size_t)c_.obs_elems * (c_.obs_kind == BB_U8 ? 1 : 4);
        Bytes o(bytes);
        uint64_t base = mix64(seed_ ^ (episode_ << 32) ^ t_);
        if (c_.obs_kind == BB_U8) {
            // cheap frame: a repeating 8-byte pattern keyed by (episode, t) -- the env must cost ~nothing
            uint64_t w = mix64(base);
            for (size_t i = 0; i + 8 <= bytes; i += 8) memcpy(o.data() + i, &w, 8);
            for (size_t i = bytes & ~(size_t)7; i < bytes; ++i) o[i] = (uint8_t)(w >> (8 * (i & 7)));
            memcpy(o.data(), &episode_, std::min<size_t>(8, bytes));
        } else {
            float* f = reinterpret_cast<float*>(o.data());
            for (uint32_t i = 0; i < c_.obs_elems; ++i)
                f[i] = (float)((int64_t)(mix64(base + i) % 2001) - 1000) * 1e-3f;
        }
        return o;
    }
    bbh_env_cfg c_;
    uint64_t seed_, episode_ = 0, t_ = 0, steps_ = 0;
};

std::unique_ptr<B200Agent> make_agent(int32_t algo, const void* cfg) {
    switch (algo) {
        case BBH_ALGO_DQN: return B200Agent::dqn(*static_cast<const bb_dqn_cfg*>(cfg));
        case BBH_ALGO_IQN: return B200Agent::iqn(*static_cast<const bb_iqn_cfg*>(cfg));
        case BBH_ALGO_SAC: return B200Agent::sac(*static_cast<const bb_sac_cfg*>(cfg));
    }
    throw Error("unknown algo");
}

size_t nz(uint64_t v) { return v == 0 ? kNever : (size_t)v; }

}  // namespace

#define BBH_BEGIN try {
#define BBH_END                                                                                 \
    return 0;                                                                                   \
    }                                                                                           \
    catch (const std::exception& e) { g_err = e.what(); return 1; }                             \
    catch (...) { g_err = "unknown C++ exception"; return 2; }

extern "C" {

const char* bbh_last_error(void) { return g_err.c_str(); }

void bbh_trainer_cfg_default(bbh_trainer_cfg* c) {  // TrainerConfig::default, trainer/config.rs:49-62
    memset(c, 0, sizeof(*c));
    c->max_opts = 0; c->opt_interval = 1; c->eval_interval = 0; c->flush_record_interval = 0;
    c->record_compute_cost_interval = 0; c->record_agent_info_interval = 0; c->warmup_period = 0; c->save_interval = 0;
    c->sync_interval = 1; c->n_actors = 1; c->n_buffer = 100;  // ActorManagerConfig::default n_buffer 100
}

int32_t bbh_train(int32_t algo, const void* agent_cfg, const bb_replay_cfg* replay_cfg, const bbh_env_cfg* env_cfg,
                  const bbh_trainer_cfg* tc, const char* save_dir, bbh_train_stat* out) {
    BBH_BEGIN
    if (!agent_cfg || !replay_cfg || !env_cfg || !tc || !out) throw Error("null argument");
    auto agent = make_agent(algo, agent_cfg);
    B200ReplayBuffer buffer(*replay_cfg);
    TrainerConfig cfg;
    cfg.max_opts = (size_t)tc->max_opts; cfg.opt_interval = std::max<size_t>(1, (size_t)tc->opt_interval);
    cfg.record_agent_info_interval = nz(tc->record_agent_info_interval); cfg.warmup_period = (size_t)tc->warmup_period;
    cfg.save_interval = nz(tc->save_interval);
    Trainer trainer(cfg);
    TrainStat st = trainer.train(std::make_unique<SyntheticEnv>(*env_cfg, tc->env_seed), SimpleStepProcessor(), *agent, buffer,
                                 save_dir ? save_dir : "");
    memset(out, 0, sizeof(*out));
    out->env_steps = st.env_steps; out->opt_steps = st.opt_steps; out->records = st.records; out->saves = st.saves;
    out->opt_seconds = st.opt_seconds; out->sample_seconds = st.sample_seconds; out->total_seconds = st.total_seconds;
    out->last_loss = st.last_loss; out->buffer_len = buffer.len();
    uint64_t n = 0;
    check(bb_agent_n_opts(agent->handle(), &n));
    out->agent_n_opts = n;
    BBH_END
}

int32_t bbh_train_offline(int32_t algo, const void* agent_cfg, bb_replay* dataset, const bbh_trainer_cfg* tc, const char* save_dir,
                          bbh_train_stat* out) {
    BBH_BEGIN
    if (!agent_cfg || !dataset || !tc || !out) throw Error("null argument");
    auto agent = make_agent(algo, agent_cfg);
    BorrowedReplayBuffer buffer(dataset);
    TrainerConfig cfg;
    cfg.max_opts = (size_t)tc->max_opts;
    cfg.record_agent_info_interval = nz(tc->record_agent_info_interval);
    cfg.save_interval = nz(tc->save_interval);
    Trainer trainer(cfg);
    TrainStat st = trainer.train_offline(*agent, buffer, save_dir ? save_dir : "");
    memset(out, 0, sizeof(*out));
    out->env_steps = st.env_steps; out->opt_steps = st.opt_steps; out->records = st.records; out->saves = st.saves;
    out->opt_seconds = st.opt_seconds; out->total_seconds = st.total_seconds;
    out->last_loss = st.last_loss; out->buffer_len = buffer.len();
    uint64_t n = 0;
    check(bb_agent_n_opts(agent->handle(), &n));
    out->agent_n_opts = n;
    BBH_END
}

int32_t bbh_train_async(int32_t algo, const void* agent_cfg, const bb_replay_cfg* replay_cfg, const bbh_env_cfg* env_cfg,
                        const bbh_trainer_cfg* tc, bbh_train_stat* out) {
    return bbh_train_async_ex(algo, agent_cfg, replay_cfg, env_cfg, tc, nullptr, nullptr, out);
}

int32_t bbh_train_async_ex(int32_t algo, const void* agent_cfg, const bb_replay_cfg* replay_cfg, const bbh_env_cfg* env_cfg,
                           const bbh_trainer_cfg* tc, bbh_learner_hook hook, void* user, bbh_train_stat* out) {
    BBH_BEGIN
    if (!agent_cfg || !replay_cfg || !env_cfg || !tc || !out) throw Error("null argument");
    B200ReplayBuffer buffer(*replay_cfg);
    AsyncTrainerConfig cfg;
    cfg.max_opts = (size_t)tc->max_opts; cfg.record_agent_info_interval = nz(tc->record_agent_info_interval);
    cfg.sync_interval = std::max<size_t>(1, (size_t)tc->sync_interval); cfg.warmup_period = (size_t)tc->warmup_period;
    bbh_env_cfg ec = *env_cfg;
    AsyncTrainStat st = train_async(
        cfg, (size_t)tc->n_actors, (size_t)tc->n_buffer, [=] { return make_agent(algo, agent_cfg); },
        [=](size_t seed) { return std::unique_ptr<Env>(new SyntheticEnv(ec, seed)); }, buffer,
        hook ? std::function<void(B200Agent&, int)>([=](B200Agent& a, int phase) { hook(a.handle(), phase, user); })
             : std::function<void(B200Agent&, int)>());
    memset(out, 0, sizeof(*out));
    out->opt_steps = cfg.max_opts; out->samples_total = st.samples_total; out->syncs = st.syncs;
    out->samples_per_sec = st.samples_per_sec; out->opt_per_sec = st.opt_per_sec; out->total_seconds = st.seconds;
    out->last_loss = st.last_loss; out->buffer_len = buffer.len();
    for (auto& a : st.actors) out->env_steps += a.env_steps;
    BBH_END
}

// Test hook: n_steps x Sampler::sample_and_push with a scripted policy (action of step i = i) and a recording buffer; row i of
// `out` = {obs episode, obs pattern word, next_obs episode, next_obs pattern word, act, reward as int, is_terminated,
// is_truncated} of the i-th pushed transition (u8 observations: bytes 0-7 hold the episode, bytes 8-15 the (episode, t) word).
int32_t bbh_sampler_trace(const bbh_env_cfg* env_cfg, uint64_t seed, uint64_t n_steps, int64_t* out) {
    BBH_BEGIN
    if (!env_cfg || !out) throw Error("null argument");
    if (env_cfg->obs_kind != BB_U8 || env_cfg->obs_elems < 16) throw Error("bbh_sampler_trace needs u8 observations of >= 16 bytes");
    struct ScriptedPolicy : Policy {
        int64_t i = 0;
        Bytes sample(const Bytes&) override { Bytes a(8); memcpy(a.data(), &i, 8); ++i; return a; }
    } policy;
    struct Recorder : ExperienceBufferBase {
        std::vector<Transition> rows;
        void push(Transition&& t) override { rows.push_back(std::move(t)); }
        size_t len() const override { return rows.size(); }
    } rec;
    Sampler sampler(std::make_unique<SyntheticEnv>(*env_cfg, seed), SimpleStepProcessor());
    for (uint64_t k = 0; k < n_steps; ++k) sampler.sample_and_push(policy, rec);
    for (size_t k = 0; k < rec.rows.size(); ++k) {
        const Transition& t = rec.rows[k];
        int64_t* o = out + 8 * k;
        memcpy(&o[0], t.obs.data(), 8); memcpy(&o[1], t.obs.data() + 8, 8);
        memcpy(&o[2], t.next_obs.data(), 8); memcpy(&o[3], t.next_obs.data() + 8, 8);
        memcpy(&o[4], t.act.data(), 8);
        o[5] = (int64_t)t.reward; o[6] = t.is_terminated; o[7] = t.is_truncated;
    }
    BBH_END
}

int32_t bbh_e2e_steps(bb_agent* agent, bb_replay* replay, const void* obs, const void* act, const void* next_obs,
                      const float* reward, const int8_t* is_terminated, const int8_t* is_truncated,
                      uint64_t obs_row_bytes, uint64_t act_row_bytes, uint64_t n_slots, uint64_t n_steps,
                      float* last_loss) {
    BBH_BEGIN
    if (!agent || !replay || !obs || !act || !next_obs || !reward || !is_terminated || !is_truncated || !n_slots)
        throw Error("null argument");
    bb_record rec;
    memset(&rec, 0, sizeof(rec));
    for (uint64_t i = 0; i < n_steps; ++i) {
        const uint64_t j = i % n_slots;
        check(bb_replay_push(replay, (const uint8_t*)obs + j * obs_row_bytes, (const uint8_t*)act + j * act_row_bytes,
                             (const uint8_t*)next_obs + j * obs_row_bytes, reward + j, is_terminated + j, is_truncated + j, 1, 0));
        check(bb_agent_opt(agent, replay, &rec));
    }
    if (last_loss) *last_loss = rec.loss;
    BBH_END
}

/* n_steps x Sampler::sample_and_push with a zero-cost environment (trainer/sampler.rs:99-144): Policy::sample on a host
 * observation (discrete action, DQN / IQN) + ExperienceBufferBase::push of the resulting host transition. */
int32_t bbh_env_steps(bb_agent* agent, bb_replay* replay, const void* obs, const void* next_obs, const float* reward,
                      const int8_t* is_terminated, const int8_t* is_truncated, uint64_t obs_row_bytes, uint64_t n_slots,
                      uint64_t n_steps, int64_t* last_act) {
    BBH_BEGIN
    if (!agent || !replay || !obs || !next_obs || !reward || !is_terminated || !is_truncated || !n_slots)
        throw Error("null argument");
    int64_t act = 0;
    for (uint64_t i = 0; i < n_steps; ++i) {
        const uint64_t j = i % n_slots;
        check(bb_agent_sample(agent, (const uint8_t*)obs + j * obs_row_bytes, 1, &act));
        check(bb_replay_push(replay, (const uint8_t*)obs + j * obs_row_bytes, &act, (const uint8_t*)next_obs + j * obs_row_bytes,
                             reward + j, is_terminated + j, is_truncated + j, 1, 0));
    }
    if (last_act) *last_act = act;
    BBH_END
}

/* The same loop through bb_actor_step: the observation crosses PCIe once, the action is chosen on the device and the
 * transition is pushed from the device-resident copies.  Zero-cost environment: step i returns obs[i % n_slots] (and, when
 * that slot's flags end the episode, next_obs[i % n_slots] plays env.reset()). */
int32_t bbh_actor_steps(bb_agent* agent, bb_replay* replay, const void* obs, const void* next_obs, const float* reward,
                        const int8_t* is_terminated, const int8_t* is_truncated, uint64_t obs_row_bytes, uint64_t n_slots,
                        uint64_t n_steps, int64_t* last_act) {
    BBH_BEGIN
    if (!agent || !replay || !obs || !next_obs || !reward || !is_terminated || !is_truncated || !n_slots)
        throw Error("null argument");
    int64_t act = 0;
    for (uint64_t i = 0; i < n_steps; ++i) {
        const uint64_t j = i % n_slots;
        const bool done = is_terminated[j] || is_truncated[j];
        check(bb_actor_step(agent, replay, (const uint8_t*)obs + j * obs_row_bytes,
                            done ? (const uint8_t*)next_obs + j * obs_row_bytes : nullptr, reward[j], is_terminated[j],
                            is_truncated[j], &act));
    }
    if (last_act) *last_act = act;
    BBH_END
}

}  // extern "C"
