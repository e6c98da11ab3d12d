#!/usr/bin/env python
"""bench.py -- grad-steps/sec of the DQN-Atari hot path (BASELINE.json configs[1]) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU (torch) path

One "step" = one Agent::opt call = replay sample+gather (uniform, ChaCha12 indices) -> Q(obs) ->
target Q(next_obs) -> MSE loss -> backward -> Adam (+ hard target copy every 10,000 steps), on a
device-resident ring of 2^20 synthetic 84x84x4 u8 transitions, batch 256, NatureCNN, 6 actions.
Prints ONE JSON line (see the keys below).  Nothing here runs under a profiler.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B = 256
N_ACT = 6
OBS_SHAPE = (4, 84, 84)
ROW = 4 * 84 * 84
WORKLOAD = "dqn_atari_pong_84x84x4_naturecnn_b256_replay1M_u8"
# algorithmic bytes of one sampled batch (u8 rows); the kernel materialises the
# batch, so the same bytes are written again.
GATHER_READ_BYTES = B * (2 * ROW + 8 + 4 + 1 + 1)
# MACs per sample: forward 9,346,048 (A=6); update = fwd + fwd_target + bwd
FWD_MACS = {"c1": 3276800, "c2": 2654208, "c3": 1806336, "l1": 1605632, "l2": 512 * N_ACT}
STEP_FLOP = 2 * B * 34107392


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle (torch fp32, all host threads) on the same workload
# ----------------------------------------------------------------------------------------------

def cpu_dqn_runner(capacity=4096, seed=42):
    """The reference's CPU path restated (oracle/): f32 TensorBatch rows + index_select gather
    (border-tch-agent/src/tensor_batch.rs:112-120), tch ops, C++-form Adam."""
    import numpy as np
    import torch
    from oracle import agent_oracle as ao
    from oracle import replay_oracle as ro
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    # the tch path stores Atari frames as f32 (border-atari-env/src/obs.rs:45-52)
    obs = torch.randint(0, 256, (capacity,) + (4, 1, 84, 84), generator=g, dtype=torch.uint8).to(torch.float32)
    next_obs = torch.roll(obs, -1, 0)
    act = torch.randint(0, N_ACT, (capacity, 1), generator=g)
    reward = torch.randint(-1, 2, (capacity,), generator=g).to(torch.float32)
    term = (torch.rand(capacity, generator=g) < 0.001).to(torch.int8)
    rng = ro.StdRng(seed)
    params = ao.atari_cnn_params(4, N_ACT, torch.Generator().manual_seed(0))
    agent = ao.DqnOracle(params, lambda p, x: ao.atari_cnn_forward(p, x), 1e-4, B, 0.99, 1.0, 10000, 1, False, None, "Mse")

    def sample():
        ix = torch.tensor([rng.next_u32() % capacity for _ in range(B)], dtype=torch.int64)
        return dict(obs=obs.index_select(0, ix), act=act.index_select(0, ix), next_obs=next_obs.index_select(0, ix),
                    reward=reward.index_select(0, ix), is_terminated=term.index_select(0, ix), ix_sample=ix)

    return lambda: agent.opt_(sample)


def time_cpu(steps, warmup, threads=None):
    import torch
    step = cpu_dqn_runner()
    if threads:
        torch.set_num_threads(threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    torch.set_num_threads(os.cpu_count() or 1)
    return steps / dt, dt


def time_cpu_env(steps, warmup):
    """env-steps/s of the reference's CPU path restated: Policy::sample (B=1 NatureCNN forward + argmax,
    dqn/base.rs:211-241) + SimpleStepProcessor::process + ExperienceBufferBase::push (base.rs:295-316) with a zero-cost
    synthetic env, as Sampler::sample_and_push (trainer/sampler.rs:99-144) does."""
    import numpy as np
    import torch
    from oracle import agent_oracle as ao
    from oracle import replay_oracle as ro
    torch.set_num_threads(os.cpu_count() or 1)
    params = ao.atari_cnn_params(4, N_ACT, torch.Generator().manual_seed(0))
    orc = ro.ReplayOracle(4096, 42, (4, 84, 84), np.uint8, (1,), np.int64)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (17,) + OBS_SHAPE, dtype=np.uint8)
    r, t, tr = np.ones(1, np.float32), np.zeros(1, np.int8), np.zeros(1, np.int8)

    def step(i):
        obs, nxt = frames[i % 16:i % 16 + 1], frames[i % 16 + 1:i % 16 + 2]
        with torch.no_grad():
            q = ao.atari_cnn_forward(params, torch.from_numpy(obs).reshape(1, 4, 1, 84, 84))
        act = np.array([[int(q.argmax(-1))]], np.int64)
        orc.push(obs, act, nxt, r, t, tr)

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    dt = time.perf_counter() - t0
    return steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, dt = time_cpu(args.steps, args.warmup)
    cores = os.cpu_count() or 1
    sample = "%d full B=256 NatureCNN update steps, replay = 4096 f32 rows (index_select gather)" % args.steps
    out = {"metric": "grad-steps/sec", "value": v, "unit": "grad-steps/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": {"workload": WORKLOAD, "batch": B, "note": "reference CPU path restated in oracle/ (Rust "
                      "toolchain absent): torch %s CPU, all host threads" % __import__("torch").__version__},
           "cpu_baseline": {"value": v, "unit": "grad-steps/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "grad-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------

def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from border_b200 import (AtariCnnConfig, Dqn, DqnConfig, DqnModelConfig, EpsilonGreedy, GenericTransitionBatch,
                             OptimizerConfig, SimpleReplayBuffer, SimpleReplayBufferConfig)
    from border_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner (and anything NCCL_DEBUG asks for) there
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = local
    # a non-default stream: the legacy default stream cannot be captured into a CUDA graph
    torch_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(torch_stream)
    stream = torch_stream.cuda_stream
    pk = peaks()

    cap = args.capacity
    rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42 + rank), device=dev)
    rb.allocate(OBS_SHAPE, np.uint8, (1,), np.int64)
    rb.fill_synthetic(cap, N_ACT, 1234 + rank)
    cfg = DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, N_ACT), opt_config=OptimizerConfig(lr=1e-4)),
                    soft_update_interval=10000, n_updates_per_opt=1, batch_size=B, discount_factor=0.99, tau=1.0,
                    train=True, explorer=EpsilonGreedy(), device=dev, critic_loss="Mse", init_seed=0)
    agent = Dqn.build(cfg)
    rb.set_stream(stream)
    agent.set_stream(stream)
    sync_mode = "none"
    if world > 1 and args.sync == "allreduce":
        from border_b200 import dist as bd
        sync_mode = bd.connect_gradient_peers(agent, dist, torch)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, repeats=1):
        """CUDA-event time of `steps` calls (max over ranks); with repeats > 1 the region is timed that many times and the
        MEDIAN is returned (20 steps are ~6 ms: one scheduler hiccup must not decide the number)."""
        for _ in range(warmup):
            fn()
        out = []
        for _ in range(repeats):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            out.append(ms)
        out.sort()
        timed.last_samples = out
        return out[len(out) // 2]

    lib = L.lib()
    n0 = __import__("ctypes").c_uint64()

    # ---- device-resident throughput: K back-to-back opt() calls
    clocks = ClockSampler(dev)
    clocks.start()
    # untimed pre-warm beyond the W warm-up steps: a fresh box ramps its clocks over more than the ~8 ms that 20 steps
    # take (a first-process run measured 9 % low without it), and the update is graph-captured after 3 eager steps
    for _ in range(300):
        agent.opt(rb)
    torch.cuda.synchronize()
    lib.bb_kernel_launch_count(None, 1)
    ms = timed(lambda: agent.opt(rb), args.steps, args.warmup, repeats=args.repeats)
    ms_samples = list(timed.last_samples)
    lib.bb_kernel_launch_count(__import__("ctypes").byref(n0), 0)
    launches_total = n0.value  # includes warm-up launches and every repeat
    clk = clocks.stop()
    launches = int(round(launches_total * args.steps / float(args.steps * args.repeats + args.warmup)))
    value = world * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers: the Trainer's inner loop in C++ (libborder_host.so,
    # the stand-in for the Rust host): push one host transition (H2D inside) + opt with record (loss D2H inside),
    # synchronous every step exactly like Agent::opt_with_record
    from border_b200 import host_loops as hl
    rng = np.random.default_rng(rank)
    n_slots = 16
    h_obs = rng.integers(0, 256, (n_slots,) + OBS_SHAPE, dtype=np.uint8)
    h_next = rng.integers(0, 256, (n_slots,) + OBS_SHAPE, dtype=np.uint8)
    h_act = rng.integers(0, N_ACT, (n_slots, 1)).astype(np.int64)
    h_rew = np.ones(n_slots, np.float32)
    h_term = np.zeros(n_slots, np.int8)
    h_trunc = np.zeros(n_slots, np.int8)
    tr = GenericTransitionBatch(h_obs[:1], h_act[:1], h_next[:1], h_rew[:1], h_term[:1], h_trunc[:1])

    def e2e_run(k):
        return hl.e2e_steps(agent, rb, h_obs, h_act, h_next, h_rew, h_term, h_trunc, k)

    e2e_steps = max(args.steps, 50)
    e2e_run(max(3, args.warmup))
    e2e_samples = []
    for _ in range(args.repeats):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host0 = time.perf_counter()
        e2e_run(e2e_steps)
        t_host1 = time.perf_counter()
        e1.record()
        barrier()
        m = max(e0.elapsed_time(e1), 1e3 * (t_host1 - t_host0))  # the loop is synchronous: device span == host span
        if world > 1:
            t = torch.tensor([m], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            m = float(t.item())
        e2e_samples.append(m)
    e2e_samples.sort()
    ms_e2e = e2e_samples[len(e2e_samples) // 2]
    e2e_value = world * e2e_steps / (ms_e2e * 1e-3)
    h2d = 2 * ROW + 8 + 4 + 1 + 1 + 2  # packed staging block of one transition (padded to 16 B)
    h2d = (h2d + 15) // 16 * 16

    # ---- env-steps/sec: Sampler::sample_and_push with a zero-cost synthetic env
    # (Policy::sample on a host obs + push), the reference's `samples_per_sec`
    env_n = max(200, args.steps)
    hl.env_steps(agent, rb, h_obs, h_next, h_rew, h_term, h_trunc, 20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hl.env_steps(agent, rb, h_obs, h_next, h_rew, h_term, h_trunc, env_n)  # C++ loop over the C ABI (bbh_env_steps)
    torch.cuda.synchronize()
    env_sps_host_push = world * env_n / (time.perf_counter() - t0)
    # ... and through bb_actor_step (the observation crosses PCIe once, explorer on the device, push from device copies)
    hl.actor_steps(agent, rb, h_obs, h_next, h_rew, h_term, h_trunc, 20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hl.actor_steps(agent, rb, h_obs, h_next, h_rew, h_term, h_trunc, env_n)
    torch.cuda.synchronize()
    env_sps = world * env_n / (time.perf_counter() - t0)
    # ... and vectorised: 8 environments per call (bb_actor_step_n; Policy::sample on a batch of n_procs observations)
    env_sps_n8 = None
    try:
        agent.actor_reset()
        o8 = np.ascontiguousarray(h_obs[:8])
        r8, z8 = np.ones(8, np.float32), np.zeros(8, np.int8)
        for _ in range(20):
            agent.actor_step_n(rb, o8, r8, z8, z8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        calls = max(100, env_n // 4)
        for _ in range(calls):
            agent.actor_step_n(rb, o8, r8, z8, z8)
        torch.cuda.synchronize()
        env_sps_n8 = world * 8 * calls / (time.perf_counter() - t0)
        agent.actor_reset()
    except Exception as e:
        env_sps_n8 = repr(e)

    # ---- roofline of the dominant kernel, measured live with CUDA events after every kernel
    prof_runs = [agent.opt_profiled(rb) for _ in range(6)][1:]
    agg = {}
    for run in prof_runs:
        for k, v in run:
            agg[k] = agg.get(k, 0.0) + v / len(prof_runs)
    step_ms_prof = sum(agg.values())
    # dominant kernel = tma_gemm_kernel (tma_gemm.cuh), 12 launches per step (c2/c3/l1 x {forward of both nets, data
    # gradient, weight gradient}): achieved = their algorithmic FLOPs / their device time, per launch = the averages.
    # The slowest and fastest single launches are reported beside it.  (At N>1 the profiled Adam also absorbs the ranks'
    # skew in its peer barrier, which is waiting, not work.)
    tcs = {k: v for k, v in agg.items() if k.split(":")[-1].startswith(("tma_gemm", "tc_gemm"))}
    if tcs:
        per = {k: roofline_for(k, v, step_ms_prof, pk) for k, v in tcs.items()}
        flop = sum(r["algorithmic_flop"] for r in per.values())
        ms_sum = sum(tcs.values())
        tf = flop / (ms_sum * 1e-3) / 1e12
        worst = min(per.values(), key=lambda r: r["achieved"])
        best = max(per.values(), key=lambda r: r["achieved"])
        traffic = [r["traffic"] for r in per.values()]
        roof = {"kernel": "tma_gemm_kernel<128xBNx32, 3xTF32, TMA-fed> (%d launches per step)" % len(tcs), "bound": "tensor",
                "achieved": tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": tf / pk["tf_sust"],
                "traffic": (sum(traffic) / len(traffic)) if all(t is not None for t in traffic) else None,
                "ms_per_launch": ms_sum / len(tcs), "launches_per_step": len(tcs), "share_of_step": ms_sum / step_ms_prof,
                "algorithmic_flop": flop / len(tcs), "peak_source": worst["peak_source"], "note": worst["note"],
                "slowest_launch": {k: worst[k] for k in ("kernel", "achieved", "frac", "ms_per_launch", "algorithmic_flop", "traffic")},
                "fastest_launch": {k: best[k] for k in ("kernel", "achieved", "frac", "ms_per_launch", "algorithmic_flop", "traffic")}}
    else:
        gemms = {k: v for k, v in agg.items() if "gemm" in k or ":tc_" in k}
        top = max((gemms or agg).items(), key=lambda kv: kv[1])
        roof = roofline_for(top[0], top[1], step_ms_prof, pk)
    # the replay kernel on its own, back to back (north-star HBM target)
    ms_g = timed(lambda: rb.batch_device(B), 2000, 20)
    us_gather = 1e3 * ms_g / 2000
    gather_gbs = 2 * GATHER_READ_BYTES / (us_gather * 1e-6) / 1e9
    roof_replay = {"kernel": "replay_sample_gather_kernel", "bound": "hbm", "achieved": gather_gbs, "peak": pk["hbm"],
                   "unit": "GB/s", "frac": gather_gbs / pk["hbm"], "traffic": NCU_TRAFFIC.get("replay_sample_gather"), "us_per_launch": us_gather,
                   "algorithmic_bytes": 2 * GATHER_READ_BYTES, "note": "read + materialised write of one B=256 batch; "
                   "frac of 8 TB/s nominal = %.3f" % (gather_gbs / 8000.0), "peak_source": pk["src"]}

    extra = None
    if world == 1 and not args.no_extra:
        try:
            extra = other_workloads(dev, stream, timed, pk)
        except Exception as e:  # the headline line must still print
            extra = {"error": repr(e)}
        try:
            # the opt-in `fast` precision mode (single-pass TF32 contractions, fp32 accumulate; include/border_b200.h:
            # bb_agent_set_precision) on the headline workload -- reported beside the line, never as its `value`
            agent.set_precision(True)
            for _ in range(20):
                agent.opt(rb)
            ms_fast = timed(lambda: agent.opt(rb), args.steps, args.warmup, repeats=3)
            agent.set_precision(False)
            for _ in range(20):
                agent.opt(rb)
            extra["dqn_fast"] = {"ms_per_step": ms_fast / args.steps, "grad_steps_per_sec": args.steps / (ms_fast * 1e-3),
                                 "precision": "single-pass TF32 operands, fp32 accumulation",
                                 "loss_error": "2-3e-4 relative on the DQN loss (tests/test_fast_mode_gpu.py)"}
        except Exception as e:
            extra["dqn_fast"] = {"error": repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt = time_cpu(args.cpu_steps, 2)
        v1, dt1 = time_cpu(max(10, args.cpu_steps // 8), 1, threads=1)   # the async example pins tch::set_num_threads(1)
        ve, dte = time_cpu_env(max(200, args.cpu_steps), 5)
        cpu = {"value": v, "unit": "grad-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "%d full B=256 NatureCNN update steps on the host CPU (torch fp32, all threads), %.1f s"
                         % (args.cpu_steps, dt),
               "one_thread": {"value": v1, "unit": "grad-steps/s", "cores": 1,
                              "sample": "%d steps with torch.set_num_threads(1) (dqn_atari_async_tch/src/main.rs:108), %.1f s" % (max(10, args.cpu_steps // 8), dt1)},
               "env_steps": {"value": ve, "unit": "env-steps/s", "cores": os.cpu_count() or 1,
                             "sample": "%d x (B=1 NatureCNN Policy::sample + replay push, zero-cost env), %.1f s" % (max(200, args.cpu_steps), dte)}}

    # ---- multi-GPU parity evidence: after every update above (different replay rows on every rank) all ranks must hold
    # bit-identical parameters -- compare a SHA-256 of the online network across ranks
    ranks_bit_identical, param_digest = None, None
    if world > 1 and sync_mode != "none":
        import hashlib
        hsh = hashlib.sha256()
        for k, v in sorted(agent.named_parameters("qnet").items()):
            hsh.update(np.ascontiguousarray(v).tobytes())
        param_digest = hsh.hexdigest()
        d = torch.tensor(list(hsh.digest()), dtype=torch.uint8, device="cuda")
        ds = [torch.empty_like(d) for _ in range(world)]
        dist.all_gather(ds, d)
        ranks_bit_identical = all(bool(torch.equal(x, ds[0])) for x in ds)

    exchange_trace = None
    if world > 1 and sync_mode != "none":
        # device-side stamps of the gradient exchange over a few updates (ns relative to the previous update's Adam start)
        import ctypes as _C
        rows, prev = [], None
        for _ in range(6):
            agent.opt(rb)
            buf = (_C.c_uint32 * 32)()
            L.check(lib.bb_agent_exchange_trace(agent.handle, buf))
            t = list(buf)
            rel = lambda a_, b_: int((a_ - b_) & 0xFFFFFFFF) if a_ and b_ else None
            rows.append({"step_ns": rel(t[16], prev) if prev else None, "fc_start_to_adam_ns": rel(t[16], t[8]),
                         "fc_wait_ns": rel(t[9], t[8]), "fc_reduce_ns": rel(t[10], t[9]), "conv_start_to_adam_ns": rel(t[16], t[12]),
                         "conv_wait_ns": rel(t[13], t[12]), "conv_reduce_ns": rel(t[14], t[13]), "fc_fence_ns": rel(t[10], t[11]), "conv_fence_ns": rel(t[14], t[15]), "adam_wait_ns": rel(t[17], t[16]), "ll_ns": rel(t[21], t[20]), "ll_start_to_adam_ns": rel(t[16], t[20]),
                         "ll_mid_ns": rel(t[25], t[24]), "ll_mid_start_to_adam_ns": rel(t[16], t[24])})
            prev = t[16]
        exchange_trace = rows[1:]

    if rank == 0:
        out = {"metric": "grad-steps/sec", "value": value, "unit": "grad-steps/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "timing": {"repeats": args.repeats, "statistic": "median", "ms_per_step_samples": [round(x / args.steps, 5) for x in ms_samples],
                          "e2e_steps": e2e_steps, "e2e_ms_per_step_samples": [round(x / e2e_steps, 5) for x in e2e_samples]},
               "config": {"workload": WORKLOAD, "batch_per_gpu": B, "replay_capacity": cap, "replay_bytes": cap * (2 * ROW + 14),
                          "n_actions": N_ACT, "optimizer": "Adam lr 1e-4", "critic_loss": "Mse",
                          "l2_policy": "inputs larger than L2 (random rows of a %.1f GB ring)" % (cap * 2 * ROW / 1e9),
                          "grad_sync": sync_mode, "unit_definition": "one B=256 update_critic (sample+gather, fwd, "
                          "target fwd, loss, bwd, Adam) per GPU"},
               "clocks": clk,
               "e2e": {"value": e2e_value, "unit": "grad-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
                       "what": "C++ host loop over the C ABI (bbh_e2e_steps): bb_replay_push(1 host transition) + bb_agent_opt(record), "
                       "loss read back every step"},
               "gpu_launches": launches,
               "env_steps_per_sec": env_sps,
               "env_steps": {"value": env_sps, "unit": "env-steps/s", "path": "bb_actor_step (device-side explorer, device->ring push)",
                             "host_sample_plus_host_push": env_sps_host_push, "n_env8_bb_actor_step_n": env_sps_n8, "h2d_bytes_per_step": ROW + 16, "d2h_bytes_per_step": 8},
               "roofline": roof, "roofline_replay": roof_replay,
               "step_flop": STEP_FLOP, "step_tflops": STEP_FLOP / (ms / args.steps * 1e-3) / 1e12,
               "kernel_breakdown_ms": {k: round(v, 5) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]}}
        if ranks_bit_identical is not None:
            out["ranks_bit_identical"] = ranks_bit_identical
            out["exchange_trace_rank0"] = exchange_trace
            out["param_sha256_rank0"] = param_digest
            out["updates_before_checksum"] = int(agent.n_opts()) if hasattr(agent, "n_opts") else None
        if extra:
            out["other_workloads"] = extra
        if cpu:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def other_workloads(dev, stream, timed, pk):
    """SURVEY.md 8(d)'s remaining measurement rows, device-resident, N=1 only (reported beside the headline,
    never as it): the gather in its bandwidth-bound regime (B scaled to 65,535 rows), the prioritized replay
    calls, the SAC (BASELINE configs[2]) and IQN (configs[3]) update steps."""
    import ctypes as C
    import numpy as np
    from border_b200 import (AtariCnnConfig, EpsilonGreedy, Iqn, IqnConfig, MlpConfig, OptimizerConfig, PerConfig, Sac,
                             SacConfig, SimpleReplayBuffer, SimpleReplayBufferConfig)
    from border_b200 import _lib as L
    lib = L.lib()
    out = {}

    def launches_of(fn):
        n = C.c_uint64()
        lib.bb_kernel_launch_count(None, 1)
        fn()
        lib.bb_kernel_launch_count(C.byref(n), 0)
        return int(n.value)

    # ---- replay sample+gather, bandwidth-bound regime: one launch moves 65,535 rows (3.7 GB read + 3.7 GB written)
    cap = 1 << 18
    rb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42), device=dev)
    rb.allocate(OBS_SHAPE, np.uint8, (1,), np.int64)
    rb.fill_synthetic(cap, N_ACT, 1234)
    rb.set_stream(stream)
    big_b = 65535  # the kernel's grid.y limit
    ms = timed(lambda: rb.batch_device(big_b), 20, 3) / 20
    nbytes = 2 * big_b * (2 * ROW + 14)
    gbs = nbytes / (ms * 1e-3) / 1e9
    out["replay_gather_b65535"] = {"kernel": "replay_sample_gather_kernel", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"],
                                   "unit": "GB/s", "frac": gbs / pk["hbm"], "frac_of_8TBs_nominal": gbs / 8000.0,
                                   "ms_per_launch": ms, "algorithmic_bytes": nbytes, "rows": big_b,
                                   "note": "same kernel as the B=256 call, batch scaled so that bandwidth (not launch "
                                           "latency) bounds it; read + materialised write", "peak_source": pk["src"]}
    rb.close()

    # ---- prioritized replay (configs[3]): sample (sum-tree descent + IS weights) and update_priority, B=256
    per = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=cap, seed=42, per_config=PerConfig(
        alpha=0.6, beta_0=0.4, beta_final=1.0, n_opts_final=500000, normalize="All")), device=dev)
    per.allocate(OBS_SHAPE, np.uint8, (1,), np.int64)
    per.fill_synthetic(cap, 4, 4321)
    per.set_stream(stream)
    us_s = 1e3 * timed(lambda: per.batch_device(B), 500, 10) / 500
    ix = np.random.default_rng(0).integers(0, cap, B).astype(np.uint64)
    td = np.random.default_rng(1).random(B).astype(np.float32)
    us_u = 1e3 * timed(lambda: per.update_priority(ix, td), 200, 10) / 200
    out["per_replay_b256"] = {"capacity": cap, "sample_gather_us": us_s, "update_priority_us_incl_h2d": us_u,
                              "note": "PER alpha 0.6, beta 0.4->1, WeightNormalizer::All; update_priority takes host ixs/td "
                                      "as the trait does"}

    # ---- IQN Atari (configs[3]): B=256, N=N'=64, AtariCnn.skip_linear features 3136, merge Mlp(3136->512->4)
    icfg = IqnConfig(f_config=AtariCnnConfig(n_stack=4, out_dim=0, skip_linear=True), m_config=MlpConfig(3136, [512], 4),
                     opt_config=OptimizerConfig(lr=1e-4), feature_dim=3136, embed_dim=64, soft_update_interval=10000,
                     batch_size=B, discount_factor=0.99, tau=1.0, train=True, sample_percents_pred="Uniform64",
                     sample_percents_tgt="Uniform64", explorer=EpsilonGreedy(), device=dev)
    iqn = Iqn.build(icfg)
    iqn.set_stream(stream)
    n_l = launches_of(lambda: iqn.opt(per))
    ms = timed(lambda: iqn.opt(per), 30, 5) / 30
    iqn_flop = 2.0 * 4 * 123.5e6 * B  # ~253 GFLOP/step
    out["iqn_atari_b256_n64"] = {"ms_per_step": ms, "grad_steps_per_sec": 1e3 / ms, "gpu_launches_per_step": n_l,
                                 "algorithmic_tflops": iqn_flop / (ms * 1e-3) / 1e12, "algorithmic_flop": iqn_flop,
                                 "frac_of_bf16_peak": iqn_flop / (ms * 1e-3) / 1e12 / pk["tf_sust"],
                                 "replay": "prioritized (sampling + IS weights; the reference's IQN never updates priorities)"}
    iqn.close()
    per.close()

    # ---- SAC Ant-like (configs[2]): obs 17, act 8, MLP[256,256], B=512, n_critics 1 (reference default) and 2
    scap = 1 << 20
    srb = SimpleReplayBuffer.build(SimpleReplayBufferConfig(capacity=scap, seed=42), device=dev)
    srb.allocate((17,), np.float32, (8,), np.float32)
    srb.fill_synthetic(scap, 0, 99)
    srb.set_stream(stream)
    for nc in (1, 2):
        scfg = SacConfig(pi_config=MlpConfig(17, [256, 256], 8), q_config=MlpConfig(25, [256, 256], 1), batch_size=512,
                         train=True, n_critics=nc, device=dev)
        sac = Sac.build(scfg)
        sac.set_stream(stream)
        n_l = launches_of(lambda: sac.opt(srb))
        ms = timed(lambda: sac.opt(srb), 200, 20) / 200
        out["sac_ant_b512_critics%d" % nc] = {"us_per_step": 1e3 * ms, "grad_steps_per_sec": 1e3 / ms,
                                               "gpu_launches_per_step": n_l, "bound": "latency (0.75 / 1.19 GFLOP per step)"}
        sac.close()
    srb.close()
    return out


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, read from the tracked ncu summary of THIS round's kernels
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` capture of one step); entries
    are used only when the kernel name recorded there is the one this build launches -- otherwise traffic is null rather
    than a stale literal."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        d = json.load(open(p))
        return d if d.get("kernel", "").startswith("tma_gemm_kernel") else {}
    except Exception:
        return {}


NCU_TRAFFIC = _ncu_traffic().get("per_layer", {})


def roofline_for(label, ms, step_ms, pk):
    """Algorithmic FLOPs of the kernel named by a profile label (phase:layer.pass:kernel)."""
    phase, layer, kernel = label.split(":")
    name = layer.split(".")[0]
    macs = FWD_MACS.get(name)
    if macs is not None and ("gemm" in kernel or kernel.startswith("tc_")):
        flop = 2.0 * B * macs  # forward, wgrad and dgrad of a layer each contract the same MACs
        tf = flop / (ms * 1e-3) / 1e12
        on_tc = kernel.startswith(("tc_", "tma_"))
        return {"kernel": label, "bound": "tensor", "achieved": tf, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": tf / pk["tf_sust"], "traffic": NCU_TRAFFIC.get(layer), "ms_per_launch": ms, "share_of_step": ms / step_ms,
                "algorithmic_flop": flop, "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside the step)",
                "note": ("TMA-fed tcgen05 kind::tf32, 3xTF32 for fp32 parity (x.[y;lo(y)] in one N=2*BN MMA from shared memory, lo(x).y from tensor memory): "
                         "algorithmic FLOPs are counted once, so the tensor pipe does 3x this; peak is the dense bf16 figure "
                         "(TF32 peak is half)")
                if on_tc else "fp32 CUDA-core implicit GEMM; peak is the dense bf16 tensor figure"}
    return {"kernel": label, "bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None,
            "traffic": None, "ms_per_launch": ms, "share_of_step": ms / step_ms, "peak_source": pk["src"]}


def run_async(args):
    """BASELINE configs[4] (border-async-trainer/src/util.rs:31-92, actor_manager/base.rs:141-175): per GPU `--actors` actor
    threads (own agent + synthetic zero-cost env, seed = actor id + rank offset) push bulks of transitions through the
    ReplayBufferProxy channel into the learner's device-resident ring while the learner runs B=256 NatureCNN updates; with
    N GPUs the learners exchange gradients every update (bb_agent_ipc_*).  One JSON line: grad-steps/s and env-steps/s of
    the whole job, measured over `--steps` synchronised updates per learner (wall clock of the slowest rank: the loop is
    host-driven, CUDA events would time the same span)."""
    import ctypes as C
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    from border_b200 import _lib as L
    from border_b200 import dist as bd
    from border_b200 import host_loops as H
    from border_b200.agents import AtariCnnConfig, DqnConfig, DqnModelConfig, EpsilonGreedy, OptimizerConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    cfg = DqnConfig(model_config=DqnModelConfig(q_config=AtariCnnConfig(4, N_ACT), opt_config=OptimizerConfig(lr=1e-4)),
                    soft_update_interval=10000, n_updates_per_opt=1, batch_size=B, discount_factor=0.99, tau=1.0, train=True,
                    explorer=EpsilonGreedy(), device=local, critic_loss="Mse", init_seed=0).to_c()
    rc = L.bb_replay_cfg()
    L.lib().bb_replay_cfg_default(C.byref(rc))
    cap = min(args.capacity, 1 << 16)
    rc.capacity, rc.seed, rc.per_config_some, rc.device = cap, 42 + rank, 0, local
    rc.obs_kind, rc.obs_elems, rc.act_kind, rc.act_elems = L.BB_U8, 4 * 84 * 84, L.BB_I64, 1
    env = H.bbh_env_cfg(L.BB_U8, 4 * 84 * 84, 1000, 0)
    warm = 4 * B
    tc = H.trainer_cfg(max_opts=args.steps + args.warmup, warmup_period=warm, sync_interval=100, n_actors=args.actors,
                       n_buffer=16, record_agent_info_interval=args.steps + args.warmup)
    state = {"sync": "none", "digest": None, "t_loop": None, "t_end": None}

    class _H:   # what border_b200.dist expects of an agent
        def __init__(self, h):
            self.handle = h

    def on_learner(handle, phase):
        if phase == 0 and world > 1:
            state["sync"] = bd.connect_gradient_peers(_H(handle), dist, torch)
        elif phase == 2:
            if world > 1:
                dist.barrier()   # every learner is warm: the updates below are synchronised across ranks anyway
            state["t_loop"] = time.perf_counter()
        elif phase == 1:
            torch.cuda.synchronize()
            state["t_end"] = time.perf_counter()
            n, no = C.c_uint64(), C.c_uint64()
            L.check(L.lib().bb_agent_model_info_size(handle, C.byref(n)))
            blob = np.empty(n.value, np.float32)   # (the size is in floats)
            L.check(L.lib().bb_agent_model_info(handle, blob.ctypes.data_as(C.c_void_p), blob.size, C.byref(no)))
            state["digest"] = hashlib.sha256(blob.tobytes()).digest()

    # zero-cost environments outrun the learner's drain: wait for room in the bounded channel instead of failing the actor
    os.environ["BBH_ACTOR_BACKPRESSURE"] = "1"
    # one emulated env step = 4 ALE frames at ~6,000 frames/s/core (the reference's frame skip, border-atari-env/src/env.rs:131)
    os.environ.setdefault("BBH_ENV_STEP_US", "600")
    t0 = time.perf_counter()
    st = H.train_async("dqn", cfg, rc, env, tc, on_learner=on_learner)
    wall = time.perf_counter() - t0
    opt_s, env_s = st["opt_per_sec"], st["samples_per_sec"]
    secs = state["t_end"] - state["t_loop"]   # the optimisation loop only (the reference's own stat also counts the warm-up)
    total_secs = st["total_seconds"]
    identical = None
    if world > 1:
        t = torch.tensor([secs, st["samples_total"] / total_secs, st["env_steps"] / total_secs], device="cuda", dtype=torch.float64)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        secs = float(mx[0].item())
        samples_ps, env_ps = float(sm[1].item()), float(sm[2].item())
        d = torch.tensor(list(state["digest"]), dtype=torch.uint8, device="cuda")
        ds = [torch.empty_like(d) for _ in range(world)]
        dist.all_gather(ds, d)
        identical = all(bool(torch.equal(x, ds[0])) for x in ds)
    else:
        samples_ps, env_ps = st["samples_total"] / total_secs, st["env_steps"] / total_secs
    if rank == 0:
        total_opts = args.steps + args.warmup
        print(json.dumps({
            "metric": "grad-steps/sec", "value": world * total_opts / secs, "unit": "grad-steps/s", "n_gpus": world,
            "steps": total_opts, "warmup": 0, "ms_per_step": 1e3 * secs / total_opts, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "async_" + WORKLOAD, "topology": "train_async: %d actor threads per GPU -> ReplayBufferProxy channel "
                       "-> learner (B=256), model sync every 100 updates; learners exchange gradients every update" % args.actors,
                       "batch_per_gpu": B, "replay_capacity": cap, "warmup_period": warm, "grad_sync": state["sync"],
                       "env": "synthetic u8 frames, %s us busy-wait per step (4 ALE frames at ~6,000 fps/core)" % os.environ["BBH_ENV_STEP_US"],
                       "timed_region": "the optimisation loop of train_async (after the %d-transition warm-up): %d updates with the "
                                       "actors pushing concurrently; wall clock of the slowest rank, device synchronised at the end" % (warm, total_opts)},
            "env_steps_per_sec": env_ps, "pushed_transitions_per_sec": samples_ps,
            "reference_stat_opt_per_sec_rank0": opt_s, "reference_stat_samples_per_sec_rank0": env_s,
            "ranks_bit_identical": identical, "wall_s_rank0": wall, "last_loss": st["last_loss"], "model_syncs": st["syncs"]}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--capacity", type=int, default=1 << 20)
    ap.add_argument("--sync", default="allreduce", choices=["allreduce", "replicas"])
    ap.add_argument("--cpu-steps", type=int, default=400)
    ap.add_argument("--repeats", type=int, default=5, help="time the K-step region this many times, report the median")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the SAC / IQN / PER / large-batch gather side measurements")
    ap.add_argument("--topology", default="sync", choices=["sync", "async"],
                    help="async = BASELINE configs[4]: actor threads per GPU feed a learner (train_async), learners "
                         "synchronised across GPUs; a side measurement, the driver's line is the default sync topology")
    ap.add_argument("--actors", type=int, default=4, help="actor threads per GPU for --topology async")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    elif args.topology == "async":
        run_async(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
