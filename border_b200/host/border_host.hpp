// border_host.hpp -- C++ mirror of the host-side loops that drive border's hot path, written
// purely against the C ABI (include/border_b200.h).  The reference is Rust and this image has no
// Rust toolchain, so this header restates -- with the reference's names, argument meaning and
// cadence -- the code that would otherwise call the ABI through bindgen:
//
//   Env::step_with_reset        border-core/src/base/env.rs:138-160
//   SimpleStepProcessor         border-core/src/generic_replay_buffer/step_proc.rs:60-137
//   Sampler::sample_and_push    border-core/src/trainer/sampler.rs:99-144
//   TrainerConfig               border-core/src/trainer/config.rs:30-88
//   Trainer::train / train_step border-core/src/trainer.rs:197-228,267-327
//   ReplayBufferProxy           border-async-trainer/src/replay_buffer_proxy.rs:52-72
//   Actor::run                  border-async-trainer/src/actor/base.rs:120-178
//   ActorManager::run           border-async-trainer/src/actor_manager/base.rs:112-186
//   AsyncTrainer::train         border-async-trainer/src/async_trainer/base.rs:204-222,299-388
//   train_async                 border-async-trainer/src/util.rs:31-92
//
// Recorders / evaluators are out of scope (host bookkeeping): Record is a
// small struct and evaluation hooks are no-ops.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "../../include/border_b200.h"

namespace border {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int32_t rc) { if (rc != 0) throw Error(bb_last_error()); }

// ----------------------------------------------------------------------------- L3 value types
using Bytes = std::vector<uint8_t>;  // one observation / action row in its replay dtype

struct Step {  // border-core/src/base/step.rs:68-94 (non-vectorised env: one row)
    Bytes act, obs;
    float reward = 0.f;
    int8_t is_terminated = 0, is_truncated = 0;
    std::optional<Bytes> init_obs;
    bool is_done() const { return is_terminated || is_truncated; }
};

struct Transition {  // GenericTransitionBatch with len() == 1 (batch.rs:89-117)
    Bytes obs, act, next_obs;
    float reward = 0.f;
    int8_t is_terminated = 0, is_truncated = 0;
};

struct Record { bb_record r{}; bool empty = true; };

// ----------------------------------------------------------------------------- traits
struct Env {  // border-core/src/base/env.rs:45-181
    virtual ~Env() = default;
    virtual Bytes reset() = 0;
    virtual Step step(const Bytes& act) = 0;
    Step step_with_reset(const Bytes& act) {  // env.rs:138-160
        Step s = step(act);
        if (s.is_done()) s.init_obs = reset();
        return s;
    }
};

struct ExperienceBufferBase {  // replay_buffer.rs:38-62
    virtual ~ExperienceBufferBase() = default;
    virtual void push(Transition&& tr) = 0;
    virtual size_t len() const = 0;
};

struct ReplayBufferBase : ExperienceBufferBase {  // replay_buffer.rs:74-127 (batch() stays on the device)
    virtual bb_replay* handle() = 0;
    // the items of one PushedItemMessage, in order (async_trainer/base.rs:280-282 pushes them one by one; an implementation
    // may move them in one call as long as the ring ends up identical)
    virtual void push_all(std::vector<Transition>&& items) { for (auto& it : items) push(std::move(it)); }
};

struct Policy {  // policy.rs:49-63
    virtual ~Policy() = default;
    virtual Bytes sample(const Bytes& obs) = 0;
};

struct Agent : Policy {  // agent.rs:24-136
    virtual void train() = 0;
    virtual void eval() = 0;
    virtual bool is_train() const = 0;
    virtual void opt(ReplayBufferBase& buffer) = 0;
    virtual Record opt_with_record(ReplayBufferBase& buffer) = 0;
    virtual void save_params(const std::string& dir) = 0;
    virtual void load_params(const std::string& dir) = 0;
};

struct ModelInfo { size_t n_opts = 0; std::vector<float> blob; };
struct SyncModel {  // border-async-trainer/src/sync_model.rs:2-13
    virtual ~SyncModel() = default;
    virtual ModelInfo model_info() = 0;
    virtual void sync_model(const ModelInfo& m) = 0;
};

// ----------------------------------------------------------------------------- B200 implementations
class B200ReplayBuffer : public ReplayBufferBase {
  public:
    explicit B200ReplayBuffer(const bb_replay_cfg& cfg) { check(bb_replay_create(&cfg, &h_)); }
    ~B200ReplayBuffer() override { bb_replay_destroy(h_); }
    B200ReplayBuffer(const B200ReplayBuffer&) = delete;
    void push(Transition&& tr) override {
        check(bb_replay_push(h_, tr.obs.data(), tr.act.data(), tr.next_obs.data(), &tr.reward, &tr.is_terminated,
                             &tr.is_truncated, 1, 0));
    }
    // one bb_replay_push of n rows (= n pushes of one row: same ring indices, same PER priorities), one H2D copy
    void push_all(std::vector<Transition>&& items) override {
        const size_t n = items.size();
        if (n == 0) return;
        if (n == 1) { push(std::move(items[0])); return; }
        const size_t ob = items[0].obs.size(), ab = items[0].act.size();
        pk_obs_.resize(n * ob); pk_next_.resize(n * ob); pk_act_.resize(n * ab);
        pk_r_.resize(n); pk_t_.resize(n); pk_tr_.resize(n);
        for (size_t i = 0; i < n; ++i) {
            const Transition& t = items[i];
            if (t.obs.size() != ob || t.next_obs.size() != ob || t.act.size() != ab) throw Error("push_all: ragged transitions");
            memcpy(pk_obs_.data() + i * ob, t.obs.data(), ob);
            memcpy(pk_next_.data() + i * ob, t.next_obs.data(), ob);
            memcpy(pk_act_.data() + i * ab, t.act.data(), ab);
            pk_r_[i] = t.reward; pk_t_[i] = t.is_terminated; pk_tr_[i] = t.is_truncated;
        }
        check(bb_replay_push(h_, pk_obs_.data(), pk_act_.data(), pk_next_.data(), pk_r_.data(), pk_t_.data(), pk_tr_.data(), n, 0));
    }
    size_t len() const override { uint64_t n = 0; check(bb_replay_len(h_, &n)); return (size_t)n; }
    bb_replay* handle() override { return h_; }
  private:
    bb_replay* h_ = nullptr;
    Bytes pk_obs_, pk_next_, pk_act_;
    std::vector<float> pk_r_;
    std::vector<int8_t> pk_t_, pk_tr_;
};

class B200Agent : public Agent, public SyncModel {
  public:
    explicit B200Agent(bb_agent* h, size_t act_bytes) : h_(h), act_bytes_(act_bytes) {}
    static std::unique_ptr<B200Agent> dqn(const bb_dqn_cfg& c) { bb_agent* h; check(bb_dqn_create(&c, &h)); return std::make_unique<B200Agent>(h, 8); }
    static std::unique_ptr<B200Agent> iqn(const bb_iqn_cfg& c) { bb_agent* h; check(bb_iqn_create(&c, &h)); return std::make_unique<B200Agent>(h, 8); }
    static std::unique_ptr<B200Agent> sac(const bb_sac_cfg& c) { bb_agent* h; check(bb_sac_create(&c, &h)); return std::make_unique<B200Agent>(h, 4 * (size_t)c.pi_config.out_dim); }
    ~B200Agent() override { bb_agent_destroy(h_); }
    B200Agent(const B200Agent&) = delete;
    Bytes sample(const Bytes& obs) override { Bytes a(act_bytes_); check(bb_agent_sample(h_, obs.data(), 1, a.data())); return a; }
    void train() override { check(bb_agent_set_train(h_, 1)); }
    void eval() override { check(bb_agent_set_train(h_, 0)); }
    bool is_train() const override { int32_t t = 0; check(bb_agent_is_train(h_, &t)); return t != 0; }
    void opt(ReplayBufferBase& b) override { check(bb_agent_opt(h_, b.handle(), nullptr)); }
    Record opt_with_record(ReplayBufferBase& b) override { Record r; check(bb_agent_opt(h_, b.handle(), &r.r)); r.empty = false; return r; }
    void save_params(const std::string& d) override { check(bb_agent_save_params(h_, d.c_str())); }
    void load_params(const std::string& d) override { check(bb_agent_load_params(h_, d.c_str())); }
    ModelInfo model_info() override {
        ModelInfo m; uint64_t n = 0, no = 0;
        check(bb_agent_model_info_size(h_, &n));
        m.blob.resize(n);
        check(bb_agent_model_info(h_, m.blob.data(), m.blob.size(), &no));
        m.n_opts = (size_t)no;
        return m;
    }
    void sync_model(const ModelInfo& m) override { check(bb_agent_sync_model(h_, m.blob.data(), m.blob.size())); }
    bb_agent* handle() { return h_; }
  private:
    bb_agent* h_;
    size_t act_bytes_;
};

// ----------------------------------------------------------------------------- sync training
class SimpleStepProcessor {  // step_proc.rs:60-137: 1-step TD transitions
  public:
    void reset(const Bytes& init_obs) { prev_obs_ = init_obs; }
    Transition process(Step&& step) {
        if (!prev_obs_) throw Error("prev_obs is not set. Forgot to call reset()?");  // step_proc.rs:104-106
        Transition t;
        t.next_obs = step.obs;
        t.obs = std::move(*prev_obs_);
        t.act = std::move(step.act);
        t.reward = step.reward;
        t.is_terminated = step.is_terminated;
        t.is_truncated = step.is_truncated;
        if (step.is_done()) {
            if (!step.init_obs) throw Error("Failed to unwrap init_obs");
            prev_obs_ = *step.init_obs;
        } else {
            prev_obs_ = std::move(step.obs);
        }
        return t;
    }
  private:
    std::optional<Bytes> prev_obs_;
};

class Sampler {  // trainer/sampler.rs:45-144
  public:
    Sampler(std::unique_ptr<Env> env, SimpleStepProcessor sp) : env_(std::move(env)), sp_(std::move(sp)) {}
    void sample_and_push(Policy& agent, ExperienceBufferBase& buffer) {
        if (!prev_obs_) {  // sampler.rs:109-115
            prev_obs_ = env_->reset();
            sp_.reset(*prev_obs_);
        }
        Bytes act = agent.sample(*prev_obs_);
        Step step = env_->step_with_reset(act);
        const bool is_done = step.is_done();
        prev_obs_ = is_done ? *step.init_obs : step.obs;   // sampler.rs:126-129
        Transition tr = sp_.process(std::move(step));
        buffer.push(std::move(tr));
        if (is_done) sp_.reset(*prev_obs_);                 // sampler.rs:138-141
    }
  private:
    std::unique_ptr<Env> env_;
    SimpleStepProcessor sp_;
    std::optional<Bytes> prev_obs_;
};

constexpr size_t kNever = std::numeric_limits<size_t>::max();

struct TrainerConfig {  // trainer/config.rs:30-88 (defaults :49-62)
    size_t max_opts = 0, opt_interval = 1, eval_interval = 0, flush_record_interval = kNever,
           record_compute_cost_interval = kNever, record_agent_info_interval = kNever, warmup_period = 0,
           save_interval = kNever;
};

struct TrainStat {  // what the reference records as average_opt_time / average_sample_time
    size_t env_steps = 0, opt_steps = 0, records = 0, saves = 0;
    double opt_seconds = 0, sample_seconds = 0, total_seconds = 0;
    float last_loss = 0.f;
};

class Trainer {  // trainer.rs:100-327
  public:
    explicit Trainer(const TrainerConfig& c) : cfg_(c) {}
    // save_dir empty => Recorder::save_model is a no-op
    TrainStat train(std::unique_ptr<Env> env, SimpleStepProcessor sp, Agent& agent, ReplayBufferBase& buffer,
                    const std::string& save_dir = "") {
        using clk = std::chrono::steady_clock;
        Sampler sampler(std::move(env), std::move(sp));
        agent.train();
        TrainStat st;
        auto t_all = clk::now();
        for (;;) {
            auto t0 = clk::now();
            sampler.sample_and_push(agent, buffer);               // trainer.rs:288
            st.sample_seconds += std::chrono::duration<double>(clk::now() - t0).count();
            st.env_steps += 1;
            bool is_opt = false;                                   // train_step, trainer.rs:197-228
            if (st.env_steps < cfg_.warmup_period) {
            } else if (st.env_steps % cfg_.opt_interval != 0) {
            } else {
                auto t1 = clk::now();
                if ((st.opt_steps + 1) % cfg_.record_agent_info_interval == 0) {
                    Record r = agent.opt_with_record(buffer);
                    st.records += 1;
                    st.last_loss = r.r.loss != 0.f ? r.r.loss : r.r.loss_critic;
                } else {
                    agent.opt(buffer);
                }
                st.opt_steps += 1;
                st.opt_seconds += std::chrono::duration<double>(clk::now() - t1).count();
                is_opt = true;
            }
            if (is_opt) {                                           // post_process, trainer.rs:231-264 (no evaluator)
                if (cfg_.save_interval > 0 && cfg_.save_interval != kNever && st.opt_steps % cfg_.save_interval == 0 &&
                    !save_dir.empty()) {
                    agent.save_params(save_dir + "/" + std::to_string(st.opt_steps));
                    st.saves += 1;
                }
            }
            if (st.opt_steps == cfg_.max_opts) break;               // trainer.rs:323-325
        }
        st.total_seconds = std::chrono::duration<double>(clk::now() - t_all).count();
        return st;
    }
    // Trainer::train_offline (trainer.rs:330-384): no environment and no sampling -- warmup_period = 0, opt_interval = 1, one
    // optimisation step per loop trip on a buffer that already holds the dataset; env_steps still counts the trips.
    TrainStat train_offline(Agent& agent, ReplayBufferBase& buffer, const std::string& save_dir = "") {
        using clk = std::chrono::steady_clock;
        cfg_.warmup_period = 0;                                     // trainer.rs:344-345
        cfg_.opt_interval = 1;
        agent.train();
        TrainStat st;
        auto t_all = clk::now();
        for (;;) {
            st.env_steps += 1;                                      // :350
            auto t1 = clk::now();
            if ((st.opt_steps + 1) % cfg_.record_agent_info_interval == 0) {   // train_step, trainer.rs:197-228
                Record r = agent.opt_with_record(buffer);
                st.records += 1;
                st.last_loss = r.r.loss != 0.f ? r.r.loss : r.r.loss_critic;
            } else {
                agent.opt(buffer);
            }
            st.opt_steps += 1;
            st.opt_seconds += std::chrono::duration<double>(clk::now() - t1).count();
            if (cfg_.save_interval > 0 && cfg_.save_interval != kNever && st.opt_steps % cfg_.save_interval == 0 &&
                !save_dir.empty()) {                                // post_process, trainer.rs:231-264 (no evaluator)
                agent.save_params(save_dir + "/" + std::to_string(st.opt_steps));
                st.saves += 1;
            }
            if (st.opt_steps == cfg_.max_opts) break;               // :380-382
        }
        st.total_seconds = std::chrono::duration<double>(clk::now() - t_all).count();
        return st;
    }
  private:
    TrainerConfig cfg_;
};

// a ReplayBufferBase over a bb_replay handle the caller owns (the offline dataset is loaded before training starts)
class BorrowedReplayBuffer : public ReplayBufferBase {
  public:
    explicit BorrowedReplayBuffer(bb_replay* h) : h_(h) {}
    void push(Transition&& tr) override {
        check(bb_replay_push(h_, tr.obs.data(), tr.act.data(), tr.next_obs.data(), &tr.reward, &tr.is_terminated, &tr.is_truncated, 1, 0));
    }
    size_t len() const override { uint64_t n = 0; check(bb_replay_len(h_, &n)); return (size_t)n; }
    bb_replay* handle() override { return h_; }
  private:
    bb_replay* h_;
};

// ----------------------------------------------------------------------------- async training
template <class T>
class Channel {  // crossbeam_channel::{bounded, unbounded} as used by util.rs:57-58, actor_manager/base.rs:137
  public:
    explicit Channel(size_t cap = 0) : cap_(cap) {}
    bool try_send(T&& v) {
        std::lock_guard<std::mutex> lk(mu_);
        if (cap_ && q_.size() >= cap_) return false;
        q_.push_back(std::move(v));
        return true;
    }
    std::vector<T> try_iter() {
        std::lock_guard<std::mutex> lk(mu_);
        std::vector<T> out(std::make_move_iterator(q_.begin()), std::make_move_iterator(q_.end()));
        q_.clear();
        return out;
    }
  private:
    size_t cap_;
    std::mutex mu_;
    std::deque<T> q_;
};

struct PushedItemMessage { size_t id; std::vector<Transition> pushed_items; };  // messages.rs

class ReplayBufferProxy : public ExperienceBufferBase {  // replay_buffer_proxy.rs:30-72
  public:
    ReplayBufferProxy(size_t id, size_t n_buffer, Channel<PushedItemMessage>& sender, const std::atomic<bool>* stop = nullptr)
        : id_(id), n_buffer_(n_buffer), sender_(sender), stop_(stop) {
        buffer_.reserve(n_buffer);
    }
    void push(Transition&& tr) override {
        buffer_.push_back(std::move(tr));
        if (buffer_.size() == n_buffer_) {
            PushedItemMessage msg{id_, std::move(buffer_)};
            buffer_ = {};
            buffer_.reserve(n_buffer_);
            // error.rs:4-7: the reference fails the actor when the bounded channel is full.  BBH_ACTOR_BACKPRESSURE=1 (the
            // benchmark's zero-cost environments produce faster than any learner drains) waits for room instead.
            static const bool backpressure = getenv("BBH_ACTOR_BACKPRESSURE") && atoi(getenv("BBH_ACTOR_BACKPRESSURE")) != 0;
            while (!sender_.try_send(std::move(msg))) {
                if (!backpressure) throw Error("SendMsgForPush");
                if (stop_ && stop_->load()) return;
                std::this_thread::yield();
            }
        }
    }
    size_t len() const override { throw Error("ReplayBufferProxy::len is unimplemented"); }
  private:
    size_t id_, n_buffer_;
    Channel<PushedItemMessage>& sender_;
    const std::atomic<bool>* stop_;
    std::vector<Transition> buffer_;
};

struct SharedModel {  // Arc<Mutex<Option<(usize, ModelInfo)>>>, actor_manager/base.rs:48,59
    std::mutex mu;
    std::optional<ModelInfo> info;
};

struct ActorStat { size_t env_steps = 0; double seconds = 0; };  // actor/stat.rs:14-23

using AgentFactory = std::function<std::unique_ptr<B200Agent>()>;
using EnvFactory = std::function<std::unique_ptr<Env>(size_t seed)>;

class Actor {  // actor/base.rs:37-178
  public:
    Actor(size_t id, AgentFactory af, EnvFactory ef, size_t env_seed, size_t n_buffer, std::atomic<bool>& stop)
        : id_(id), af_(std::move(af)), ef_(std::move(ef)), env_seed_(env_seed), n_buffer_(n_buffer), stop_(stop) {}
    ActorStat run(Channel<PushedItemMessage>& sender, SharedModel& model_info) {
        auto agent = af_();
        ReplayBufferProxy buffer(id_, n_buffer_, sender, &stop_);
        Sampler sampler(ef_(env_seed_), SimpleStepProcessor());
        size_t n_opt_steps = 0;
        auto t0 = std::chrono::steady_clock::now();
        for (;;) {  // sync_model_first: wait for the initial model (actor/base.rs:72-92)
            std::lock_guard<std::mutex> lk(model_info.mu);
            if (model_info.info) { agent->sync_model(*model_info.info); n_opt_steps = model_info.info->n_opts; break; }
            std::this_thread::yield();
        }
        agent->train();
        ActorStat st;
        for (;;) {
            {  // sync_model if newer (actor/base.rs:94-118)
                std::lock_guard<std::mutex> lk(model_info.mu);
                if (model_info.info && model_info.info->n_opts > n_opt_steps) {
                    agent->sync_model(*model_info.info);
                    n_opt_steps = model_info.info->n_opts;
                }
            }
            sampler.sample_and_push(*agent, buffer);
            st.env_steps += 1;
            if (stop_.load()) break;
        }
        st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return st;
    }
  private:
    size_t id_;
    AgentFactory af_;
    EnvFactory ef_;
    size_t env_seed_, n_buffer_;
    std::atomic<bool>& stop_;
};

struct AsyncTrainerConfig {  // async_trainer/config.rs:11-28
    size_t max_opts = 0, eval_interval = kNever, flush_record_interval = kNever, record_compute_cost_interval = kNever,
           record_agent_info_interval = kNever, save_interval = kNever, sync_interval = 1, warmup_period = 0;
};

struct AsyncTrainStat { double samples_per_sec = 0, opt_per_sec = 0, seconds = 0; size_t samples_total = 0, syncs = 0; float last_loss = 0.f;
                        std::vector<ActorStat> actors; };

// train_async (util.rs:31-92): N actor threads -> channel -> learner thread (the caller's).
// on_learner(agent, phase): called on the learner thread with phase 0 right after the learner agent exists (a data-parallel
// job connects its gradient peers there), phase 2 when the warm-up is over and the optimisation loop starts, phase 1 after
// the last optimisation step (parameter checksums).
inline AsyncTrainStat train_async(const AsyncTrainerConfig& cfg, size_t n_actors, size_t n_buffer, AgentFactory af,
                                  EnvFactory ef, ReplayBufferBase& buffer,
                                  const std::function<void(B200Agent&, int)>& on_learner = nullptr) {
    using clk = std::chrono::steady_clock;
    Channel<PushedItemMessage> items(1000 * std::max<size_t>(1, n_actors));  // bounded(1000) per forwarding hop
    SharedModel shared;
    std::atomic<bool> stop{false};
    std::vector<ActorStat> stats(n_actors);
    std::vector<std::thread> threads;
    std::vector<std::string> errors(n_actors);
    for (size_t i = 0; i < n_actors; ++i)  // ActorManager::run, seed = actor id (actor_manager/base.rs:141-175)
        threads.emplace_back([&, i] {
            try {
                Actor a(i, af, ef, i, n_buffer, stop);
                stats[i] = a.run(items, shared);
            } catch (const std::exception& e) { errors[i] = e.what(); }
        });
    AsyncTrainStat out;
    try {
        auto agent = af();                                         // async_trainer/base.rs:314
        agent->train();
        if (on_learner) on_learner(*agent, 0);
        auto sync = [&] {                                          // :268-272
            ModelInfo m = agent->model_info();
            std::lock_guard<std::mutex> lk(shared.mu);
            shared.info = std::move(m);
            out.syncs += 1;
        };
        auto update_replay_buffer = [&] {                          // :275-284
            for (auto& msg : items.try_iter()) {
                out.samples_total += msg.pushed_items.size();
                buffer.push_all(std::move(msg.pushed_items));
            }
        };
        auto t_all = clk::now();
        sync();                                                    // :325
        while (buffer.len() < cfg.warmup_period) {                 // :328-334
            update_replay_buffer();
            for (auto& e : errors) if (!e.empty()) throw Error("actor failed: " + e);
            std::this_thread::yield();
        }
        if (on_learner) on_learner(*agent, 2);                     // warm-up over: the optimisation loop starts
        size_t opt_steps = 0;
        for (;;) {
            update_replay_buffer();                                // :340
            if (buffer.len() >= cfg.warmup_period) {               // train_step, :204-222
                if ((opt_steps + 1) % cfg.record_agent_info_interval == 0) {
                    Record r = agent->opt_with_record(buffer);
                    out.last_loss = r.r.loss != 0.f ? r.r.loss : r.r.loss_critic;
                } else {
                    agent->opt(buffer);
                }
                opt_steps += 1;
            }
            if (opt_steps % cfg.sync_interval == 0) sync();        // post_process, :258-261
            if (opt_steps == cfg.max_opts) {                       // :369-375
                stop.store(true);
                items.try_iter();
                sync();
                if (on_learner) on_learner(*agent, 1);
                break;
            }
        }
        out.seconds = std::chrono::duration<double>(clk::now() - t_all).count();
        out.samples_per_sec = out.samples_total / out.seconds;
        out.opt_per_sec = cfg.max_opts / out.seconds;
    } catch (...) {
        stop.store(true);
        { std::lock_guard<std::mutex> lk(shared.mu); if (!shared.info) shared.info = ModelInfo{}; }
        for (auto& t : threads) t.join();
        throw;
    }
    for (auto& t : threads) t.join();
    for (auto& e : errors) if (!e.empty()) throw Error("actor failed: " + e);
    out.actors = stats;
    return out;
}

}  // namespace border
