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


def time_cpu(steps, warmup):
    step = cpu_dqn_runner()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
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
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (no version banner)
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

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
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
        return ms

    lib = L.lib()
    n0 = __import__("ctypes").c_uint64()

    # ---- device-resident throughput: K back-to-back opt() calls
    clocks = ClockSampler(dev)
    clocks.start()
    lib.bb_kernel_launch_count(None, 1)
    ms = timed(lambda: agent.opt(rb), args.steps, args.warmup)
    lib.bb_kernel_launch_count(__import__("ctypes").byref(n0), 0)
    launches_total = n0.value  # includes warm-up launches
    clk = clocks.stop()
    launches = int(round(launches_total * args.steps / float(args.steps + args.warmup)))
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

    e2e_steps = max(10, args.steps // 2)
    e2e_run(max(3, args.warmup // 2))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    e2e_run(e2e_steps)
    t_host1 = time.perf_counter()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e3 * (t_host1 - t_host0))  # the loop is synchronous: device span == host span
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * e2e_steps / (ms_e2e * 1e-3)
    h2d = 2 * ROW + 8 + 4 + 1 + 1 + 2  # packed staging block of one transition (padded to 16 B)
    h2d = (h2d + 15) // 16 * 16

    # ---- env-steps/sec: Sampler::sample_and_push with a zero-cost synthetic env
    # (Policy::sample on a host obs + push), the reference's `samples_per_sec`
    def env_step():
        agent.sample(h_obs[:1])
        rb.push(tr)

    env_n = max(50, args.steps)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(env_n):
        env_step()
    torch.cuda.synchronize()
    env_sps = world * env_n / (time.perf_counter() - t0)

    # ---- roofline of the dominant kernel, measured live with CUDA events after every kernel
    prof_runs = [agent.opt_profiled(rb) for _ in range(6)][1:]
    agg = {}
    for run in prof_runs:
        for k, v in run:
            agg[k] = agg.get(k, 0.0) + v / len(prof_runs)
    step_ms_prof = sum(agg.values())
    top = max(agg.items(), key=lambda kv: kv[1])
    roof = roofline_for(top[0], top[1], step_ms_prof, pk)
    # the replay kernel on its own, back to back (north-star HBM target)
    ms_g = timed(lambda: rb.batch_device(B), 2000, 20)
    us_gather = 1e3 * ms_g / 2000
    gather_gbs = 2 * GATHER_READ_BYTES / (us_gather * 1e-6) / 1e9
    roof_replay = {"kernel": "replay_sample_gather_kernel", "bound": "hbm", "achieved": gather_gbs, "peak": pk["hbm"],
                   "unit": "GB/s", "frac": gather_gbs / pk["hbm"], "traffic": None, "us_per_launch": us_gather,
                   "algorithmic_bytes": 2 * GATHER_READ_BYTES, "note": "read + materialised write of one B=256 batch; "
                   "frac of 8 TB/s nominal = %.3f" % (gather_gbs / 8000.0), "peak_source": pk["src"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt = time_cpu(args.cpu_steps, 2)
        cpu = {"value": v, "unit": "grad-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "%d full B=256 NatureCNN update steps on the host CPU (torch fp32, all threads), %.1f s"
                         % (args.cpu_steps, dt)}

    if rank == 0:
        out = {"metric": "grad-steps/sec", "value": value, "unit": "grad-steps/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
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
               "roofline": roof, "roofline_replay": roof_replay,
               "step_flop": STEP_FLOP, "step_tflops": STEP_FLOP / (ms / args.steps * 1e-3) / 1e12,
               "kernel_breakdown_ms": {k: round(v, 5) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]}}
        if cpu:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def roofline_for(label, ms, step_ms, pk):
    """Algorithmic FLOPs of the kernel named by a profile label (phase:layer.pass:kernel)."""
    phase, layer, kernel = label.split(":")
    name = layer.split(".")[0]
    macs = FWD_MACS.get(name)
    if macs is not None and ("gemm" in kernel or kernel.startswith("tc_")):
        flop = 2.0 * B * macs  # forward, wgrad and dgrad of a layer each contract the same MACs
        tf = flop / (ms * 1e-3) / 1e12
        on_tc = kernel.startswith("tc_")
        return {"kernel": label, "bound": "tensor", "achieved": tf, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": tf / pk["tf_sust"], "traffic": None, "ms_per_launch": ms, "share_of_step": ms / step_ms,
                "algorithmic_flop": flop, "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside the step)",
                "note": ("tcgen05 kind::tf32, 3 MMA passes per product (3xTF32 for fp32 parity): algorithmic FLOPs are "
                         "counted once, so the tensor pipe does 3x this; peak is the dense bf16 figure (TF32 peak is half)")
                if on_tc else "fp32 CUDA-core implicit GEMM; peak is the dense bf16 tensor figure"}
    return {"kernel": label, "bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None,
            "traffic": None, "ms_per_launch": ms, "share_of_step": ms / step_ms, "peak_source": pk["src"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--capacity", type=int, default=1 << 20)
    ap.add_argument("--sync", default="allreduce", choices=["allreduce", "replicas"])
    ap.add_argument("--cpu-steps", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
