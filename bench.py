#!/usr/bin/env python
"""bench.py -- DCMoE layer forward throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c2|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward of one DCMoE layer (router -> plan -> permute -> grouped FFN -> combine) over one
batch of synthetic activations.
  N = 1 : BASELINE.json configs[1] -- batch 8 x 2048 tokens, bf16, utils/config.json dims  (--workload c2).
  N > 1 : BASELINE.json configs[3] -- expert-parallel, global batch 64 x 4096 tokens sharded by sequence,
          8/N routed experts per rank, gate + shared experts replicated ("strong" scaling: the global batch
          is fixed)  (--workload c4; `--gpus 1 --workload c4` runs the same global batch on one GPU, the
          same-workload reference point of the scaling curve).
Prints ONE JSON line (rank 0).  `value` = tokens/s with inputs resident in HBM (CUDA events, max over ranks);
`e2e` = the same through the public module call with pinned HOST buffers (H2D of the step's input and D2H of
the step's outputs inside the timed region); `roofline` describes the dominant kernel (GEMM-1 of the grouped
FFN) from CUDA events recorded between the kernel launches of the timed steps; `cpu_baseline` is the reference
path timed on this box's host cores (the unmodified reference block when its tree is present, else the oracle
port) on BASELINE.json configs[0] and on a bounded sample of the same workload.

`parity` is a gate, not a statistic: after the timed loop the last step's outputs are checked -- at N = 1 against the
CPU oracle on ALL rows (routing bit-exact given the logits, outputs within rtol 1e-2), at N > 1 additionally against a
single-GPU forward of the concatenated batch (bit-equal) -- and a failure makes the run exit non-zero.  The inputs
rotate, so the checked step does not repeat its predecessor's data.

--impl reference times the reference's CPU implementation of the path on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dcmoe_layer_tokens_per_sec"
UNIT = "tokens/s"
KERNELS_PER_STEP = 6  # router, plan(+aux), permute, ffn gemm-1, ffn gemm-2, combine
FLOP_PER_ROW_GEMM1 = 4 * 2048 * 2752  # gate + up projections
FLOP_PER_ROW_GEMM2 = 2 * 2048 * 2752
LAYER_WEIGHT_BYTES_ROUTED = 8 * 3 * 2048 * 2752 * 2


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"], tflops_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (a thread polling every
    20 ms: NVML queries take driver locks, so polling faster -- or spawning nvidia-smi -- stalls kernel launches)."""

    def __init__(self, gpu_index: int):
        self.gpu_index, self.samples, self._stop, self.thread, self.err = gpu_index, [], False, None, None
        self.t_begin = self.t_end = None

    def start(self):
        if os.environ.get("DCMOE_BENCH_NO_NVML"):      # diagnosis of launch stalls: run without the NVML thread
            self.err = "disabled by DCMOE_BENCH_NO_NVML"
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # noqa: BLE001
            self.err = str(exc)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), float(sm), int(rs), 0.0))
            except Exception as exc:  # noqa: BLE001
                self.err = str(exc)
                return
            time.sleep(0.02)

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        self._stop = True
        if self.thread is not None:
            self.thread.join(timeout=1)
        if self.err is not None and not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[f"nvml unavailable: {self.err}"])
        nv = self.nv
        inside = [s for s in self.samples if self.t_begin is not None and self.t_begin <= s[0] <= (self.t_end or 1e30)]
        use = inside or self.samples
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(s[2] & bit for s in use))
        sm = [s[1] for s in use]
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_min_mhz=min(sm) if sm else None, sm_max_mhz=self.sm_max,
                    reasons=reasons, samples=len(use),
                    samples_in_timed_region=len(inside))


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def bind_rank_to_cores(local_rank: int, world: int):
    """Give every rank its own slice of the host cores the GPU is attached to (NVML's ideal CPU set; all cores if
    unknown) BEFORE it allocates pinned memory: first touch then places the rank's host buffers on the GPU's NUMA node
    and the ranks' copy / launch threads do not migrate over each other.  Returns a description for the JSON line."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(cores) // 64) + 1)
            ideal = [c for c in cores if (words[c // 64] >> (c % 64)) & 1]
            if ideal:
                cores = ideal
        except Exception:  # noqa: BLE001
            pass
        if world > 1 and len(cores) >= 2 * world:
            per = len(cores) // world
            mine = cores[local_rank * per:(local_rank + 1) * per]
        else:
            mine = cores
        os.sched_setaffinity(0, mine)
        return {"cores": f"{mine[0]}-{mine[-1]}", "n": len(mine)}
    except Exception as exc:  # noqa: BLE001
        return {"error": str(exc)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on the host cores
def _time_calls(fn, reps: int, warmup: int):
    for _ in range(warmup):
        fn()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return times


def cpu_reference_runner(tokens: int, dtype_name: str):
    """Returns (callable running one forward over 1 x tokens, kind).  kind = "reference": the UNMODIFIED reference block
    (oracle/ref_loader.py imports it from DCMOE_REFERENCE_ROOT, default /root/reference -- present in the build
    container only); "port": the CPU restatement oracle/dcmoe_oracle.py."""
    import torch

    from oracle import dcmoe_oracle as O
    from oracle import ref_loader

    dt = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    x = torch.randn(1, tokens, 2048, generator=torch.Generator().manual_seed(1235)).to(dt)
    if ref_loader.reference_available():
        block = ref_loader.build_reference_block(dtype=dt, seed=0)
        with torch.no_grad():
            return (lambda: block(x, None, None)), "reference"
    W = O.make_weights(seed=0, dtype=dt)
    return (lambda: O.forward(x, W)), "port"


def cpu_arm(sample_tokens: int, reps: int, warmup: int = 1):
    """configs[0] (1 x 512, fp32 and bf16) + a bounded bf16 sample of the benched workload."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {"cores": cores, "unit": UNIT}
    with torch.no_grad():
        for name, tokens, dn in (("config1_fp32", 512, "fp32"), ("config1_bf16", 512, "bf16")):
            fn, kind = cpu_reference_runner(tokens, dn)
            t = _time_calls(fn, reps=3, warmup=1)
            out[name] = {"tokens_per_s_best": tokens / min(t), "tokens_per_s_median": tokens / statistics.median(t),
                         "ms_best": 1e3 * min(t), "kind": kind}
        fn, kind = cpu_reference_runner(sample_tokens, "bf16")
        times = _time_calls(fn, reps=reps, warmup=warmup)
    out["kind"] = kind
    out["sample_times_s"] = times
    out["value"] = sample_tokens / (sum(times) / len(times))
    out["best"] = sample_tokens / min(times)
    return out


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    sample = 2048
    arm = cpu_arm(sample, reps=args.steps, warmup=max(1, min(args.warmup, 2)))
    times = arm.pop("sample_times_s")
    ms = 1e3 * sum(times) / len(times)
    workload = workload_name(args)
    desc = {"c2": "configs[1]: single DCMoE layer bf16, batch 8 x 2048 tokens",
            "c4": "configs[3]: expert-parallel DCMoE, global batch 64 x 4096 tokens"}[workload]
    what = ("the unmodified reference block (UniMoEAudioSparseMoeBlock, imported by oracle/ref_loader.py)" if arm["kind"] == "reference"
            else "CPU oracle port of the reference path (the reference tree is not on this box)")
    line = {
        "impl": "reference", "metric": METRIC, "value": arm["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak" if workload == "c2" else "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": desc, "note": f"{what}; each step = one forward over a bounded sample of {sample} tokens of that workload"},
        "cpu_baseline": {"value": arm["value"], "unit": UNIT, "cores": arm["cores"], "kind": arm["kind"],
                         "sample": f"1 x {sample} tokens, bf16, all host threads, mean of {len(times)} steps",
                         "config1_fp32": arm["config1_fp32"], "config1_bf16": arm["config1_bf16"]},
        "e2e": {"value": arm["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def workload_name(args) -> str:
    if args.workload != "auto":
        return args.workload
    return "c2" if args.gpus == 1 else "c4"


def build_module(dev, dt):
    import torch

    from unimoe_audio_b200 import DCMoE

    cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
               mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
               shared_intermediate_size=1376, router_jitter_noise=0.01, hidden_act="silu", ep_size=1)
    with torch.device("meta"):
        m = DCMoE(cfg)
    m = m.to(dt).to_empty(device=dev).eval()
    gen = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for _, p in sorted(m.named_parameters(), key=lambda kv: kv[0]):   # N(0, 0.02^2): initializer_range
            p.copy_((torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(dt))
    return m


def oracle_parity(m, x, out, max_rows=None):
    """Check one forward's 6-tuple against the CPU oracle (oracle/dcmoe_oracle.py, pinned to the unmodified reference
    block by tests/golden/): routing outputs bit-exact given the GPU's logits, logits within one bf16 ulp-scale of the
    CPU gate projection, layer output |a-b| <= rtol*|b| + rtol*max|b| with rtol 1e-2, on the first `max_rows` rows (all
    rows when None).  The routing of a row depends on that row only, so a prefix is a valid sub-problem -- except for
    the aux loss, which is only compared when every row is checked."""
    import torch

    from oracle import dcmoe_oracle as O

    n = x.shape[0] * x.shape[1] if max_rows is None else min(max_rows, x.shape[0] * x.shape[1])
    W = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    xc = x.reshape(-1, x.shape[-1])[:n].cpu().unsqueeze(0)
    lg = out[1][:n].cpu()
    ref = O.forward(xc, W, None, logits=lg)
    own_logits = torch.nn.functional.linear(xc[0], W["gate.weight"])
    rtol = 1e-2
    a, b = out[0].reshape(-1, x.shape[-1])[:n].float().cpu(), ref.final_hidden_states[0].float()
    bound = rtol * b.abs() + rtol * b.abs().max()
    res = {
        "rows_checked": int(n),
        "routing_bit_exact": bool(torch.equal(out[2][:n].cpu(), ref.dynamic_top_k) and torch.equal(out[3][:n].cpu(), ref.expert_mask)
                                  and torch.equal(out[4][:n].cpu(), ref.global_weight)),
        "logits_max_abs_diff_vs_cpu_gate": float((lg.float() - own_logits.float()).abs().max()),
        "max_rel": float(((a - b).abs() / (b.abs() + b.abs().max())).max()),
        "rel_fro": float((a - b).norm() / b.norm()),
        "output_within_rtol_1e-2": bool(((a - b).abs() <= bound).all()),
    }
    if n == x.shape[0] * x.shape[1]:
        res["aux_loss"] = [float(out[5]), float(ref.aux_loss)]
        res["aux_within_2e-3"] = bool(abs(float(out[5]) - float(ref.aux_loss)) <= 2e-3 * max(1.0, abs(float(ref.aux_loss))))
    res["ok"] = bool(res["routing_bit_exact"] and res["output_within_rtol_1e-2"] and res["logits_max_abs_diff_vs_cpu_gate"] <= 0.0625
                     and res.get("aux_within_2e-3", True))
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    binding = bind_rank_to_cores(local_rank, world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dt = torch.bfloat16
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    workload_id = workload_name(args)

    m = build_module(dev, dt)
    if workload_id == "c2":
        if world != 1:
            raise SystemExit("--workload c2 (configs[1]) is the single-GPU configuration")
        B, S = 8, 2048
        workload = "configs[1]: single DCMoE layer bf16 on 1xB200, batch 8 x 2048 tokens, utils/config.json dims"
        scaling = "weak"
    else:
        Bg, S = 64, 4096
        B = Bg // world
        workload = (f"configs[3]: expert-parallel DCMoE across {world} B200, global batch 64 x 4096 tokens "
                    f"({B} sequences and {8 // world} routed experts per rank)" if world > 1 else
                    "configs[3] on ONE B200 (the same-workload point of the scaling curve): batch 64 x 4096 tokens, all 8 routed experts")
        scaling = "strong"
    if world == 1:
        layer = m
        parallelism = "single GPU"
    else:
        from unimoe_audio_b200.ep import ExpertParallelDCMoE
        layer = ExpertParallelDCMoE(m, dist.group.WORLD)
        parallelism = f"ep{world}"
    T = B * S
    n_rot = 4  # rotate inputs: 4 x (>= 64 MiB) activations + 304 MB weights per step >> 126 MB L2
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = [torch.randn(B, S, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt) for _ in range(n_rot)]

    # ---- stage events (recorded between kernel launches on the launching stream) ----
    stage_events = []

    def hook(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        stage_events[-1].append((name, ev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    out = None
    for i in range(args.warmup):
        # `out` is kept alive across iterations exactly as in the timed loop: two output sets are live while a step
        # is enqueued, so the caching allocator's second output block (a cudaMalloc, 2-200 ms) is paid for here
        out = layer(xs[i % n_rot], None, None)
    barrier()

    m.stage_hook = hook
    for i in range(2):           # untimed steps with the event hooks on (first-use costs of the hook path stay outside)
        stage_events.append([])
        out = layer(xs[i % n_rot], None, None)
    barrier()
    stage_events.clear()
    if world > 1 and hasattr(layer, "comm_events"):
        layer.comm_events.clear()
    if rank == 0:
        sampler.mark_begin()
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    import gc
    gc.collect()
    gc.disable()                 # a collection pause in the enqueue loop starves the GPU for several steps
    host_t0 = time.perf_counter()
    last_idx = 0
    for i in range(args.steps):
        stage_events.append([])
        s, e = step_ev[i]
        s.record()
        last_idx = (args.warmup + i) % n_rot
        out = layer(xs[last_idx], None, None)
        e.record()
    host_ms_per_step = (time.perf_counter() - host_t0) * 1e3 / args.steps   # enqueue time only (no sync)
    gc.enable()
    barrier()
    if rank == 0:
        sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    m.stage_hook = None
    my_ms = step_ev[0][0].elapsed_time(step_ev[-1][1])     # K steps back to back, first start -> last end
    total_ms, rank_ms = my_ms, [my_ms]
    if world > 1:
        t = torch.tensor([my_ms], device=dev, dtype=torch.float64)
        allt = torch.empty(world, device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allt, t)
        rank_ms = allt.tolist()
        total_ms = max(rank_ms)
    ms_per_step = total_ms / args.steps
    value = T * world / (ms_per_step * 1e-3)

    # per-stage durations (median over the timed steps)
    stages = {}
    for evs in stage_events:
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            stages.setdefault(n1, []).append(e0.elapsed_time(e1))
    stage_ms = {k: statistics.median(v) for k, v in stages.items()}   # median: robust to a host hiccup in one step
    comm_ms = {}
    if world > 1 and getattr(layer, "comm_events", None):
        acc = {}
        for name, a, b in layer.comm_events:
            acc.setdefault(name, []).append(a.elapsed_time(b))
        comm_ms = {k: statistics.median(v) for k, v in acc.items()}
    ep_path = getattr(layer, "last_path", None)

    # routed rows of the last step (A = sum_t r_t) -> algorithmic FLOPs of the grouped GEMMs
    ws = m.last_workspace
    n_rows_local = int(ws.mtiles[: int(ws.n_mtiles.item()), 3].sum().item())  # valid rows in this rank's row space
    roofline = None
    g1_key = "ffn_gemm1" if "ffn_gemm1" in stage_ms else ("ffn_gemm1_routed" if "ffn_gemm1_routed" in stage_ms else None)
    if g1_key is not None:
        # overlapped token-dispatch expert parallelism launches GEMM-1 twice (shared tiles under the dispatch, then the
        # routed tiles): the roofline line then describes the routed launch, which runs alone on the GPU
        g1_rows = n_rows_local if g1_key == "ffn_gemm1" else n_rows_local - T
        flops = g1_rows * FLOP_PER_ROW_GEMM1
        ach = flops / (stage_ms[g1_key] * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath) and workload_id == "c2" and world == 1:
            try:
                tj = json.load(open(tpath))
                traffic, traffic_src = tj.get("ffn_gemm1_dram_bytes_per_launch"), tj.get("source")
            except Exception:  # noqa: BLE001
                traffic = None
        # the step is a few ms and the timed region tens of ms: the kernel runs in the burst regime, so the
        # denominator is the BURST cuBLAS figure (the larger, i.e. stricter, of the two measured peaks)
        roofline = {"kernel": "ffn_gemm_kernel<SwiGLU> (grouped GEMM-1, tcgen05)", "bound": "tensor", "achieved": ach,
                    "peak": peaks["tflops_burst"], "unit": "TFLOP/s", "frac": ach / peaks["tflops_burst"],
                    "peak_source": f"bf16_tflops (burst) of {peaks['source']}; frac_of_sustained uses bf16_tflops_sustained",
                    "frac_of_sustained": ach / peaks["tflops_sustained"],
                    "traffic": traffic, "traffic_source": traffic_src or "not captured for this configuration (ncu --set full is a separate run)",
                    "rows_per_launch": g1_rows,
                    "flop_per_row": FLOP_PER_ROW_GEMM1, "avg_launch_ms": stage_ms[g1_key], "launch": g1_key}
        # the HBM-bound kernels, same events (algorithmic bytes per SURVEY.md 8d / BASELINE.md section 4)
        A = n_rows_local - T
        hbm = {"router": T * 4192, "permute": (T + A) * 4096, "combine": (A + T) * 4096 + T * 4096}
        roofline["hbm_kernels"] = {
            k: {"achieved_gbs": b / (stage_ms[k2] * 1e-3) / 1e9, "frac": b / (stage_ms[k2] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "bytes": b, "ms": stage_ms[k2]}
            for k, k2, b in (("router", "router", hbm["router"]),
                             ("permute", "permute" if "permute" in stage_ms else "ep_dispatch", hbm["permute"]),
                             ("combine", "combine" if "combine" in stage_ms else "ep_combine", hbm["combine"]))
            if k2 in stage_ms}
        g2_key = "ffn_gemm2" if "ffn_gemm2" in stage_ms else "ffn_gemm2_routed"
        if g2_key in stage_ms:
            roofline["gemm2"] = {"achieved": g1_rows * FLOP_PER_ROW_GEMM2 / (stage_ms[g2_key] * 1e-3) / 1e12,
                                 "unit": "TFLOP/s", "ms": stage_ms[g2_key], "launch": g2_key}

    if world > 1 and roofline is not None:
        nv = {}
        if ep_path == "gather":
            # weight-gather expert parallelism: what crosses NVLink per step is the remote experts' packed weights
            remote_bytes = LAYER_WEIGHT_BYTES_ROUTED * (world - 1) // world
            ms = comm_ms.get("weight_fetch")
            if ms:
                gbs = remote_bytes / (ms * 1e-3) / 1e9
                nv["weight_fetch"] = {"remote_weight_bytes": remote_bytes, "ms": ms, "achieved_gbs": gbs, "frac_of_770": gbs / 770.0,
                                      "exposed_wait_ms": stage_ms.get("wait_weights"),
                                      "note": "copy-engine peer copies on a side stream, double buffered: the fetch of step i+1 "
                                              "runs under the GEMMs of step i; exposed_wait_ms is what the compute stream waited"}
        else:
            # token dispatch: rows that leave this rank (dispatch) / come back (combine gather), 4096 B each, against the
            # measured peer-copy bandwidth of this pool (B200_PROFILING.md: 770 GB/s per direction per GPU)
            A_sent = int(out[3][:, :8].sum().item())
            remote = A_sent * (world - 1) / world
            for key, name in (("ep_dispatch", "dispatch"), ("ep_combine_gather", "combine_gather"), ("ep_combine", "combine")):
                ms = comm_ms.get(key, stage_ms.get(key))
                if ms:
                    gbs = remote * 4096 / (ms * 1e-3) / 1e9
                    nv[name] = {"remote_rows_estimate": int(remote), "ms": ms, "achieved_gbs": gbs, "frac_of_770": gbs / 770.0}
        roofline["nvlink"] = nv

    # ---- parity gates on the LAST timed step (its input differs from the previous step's) ----
    x_last = xs[last_idx]
    parity = {}
    if world == 1:
        parity["oracle"] = oracle_parity(m, x_last, out, None if T <= 65536 else 32768)
        parity["routing_bit_exact"] = parity["oracle"]["routing_bit_exact"]
        parity["max_rel"] = parity["oracle"]["max_rel"]
        parity["ok"] = parity["oracle"]["ok"]
    else:
        # every rank's step output, gathered, must equal ONE single-GPU forward over the concatenated batch
        x_all = torch.empty((world * B, S, 2048), dtype=dt, device=dev)
        dist.all_gather_into_tensor(x_all, x_last.contiguous())
        gathered = []
        for i in (0, 1, 2, 3, 4):
            t = out[i].contiguous()
            g = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            dist.all_gather_into_tensor(g, t)
            gathered.append(g)
        verdict = torch.zeros(2, dtype=torch.int32, device=dev)
        if rank == 0:
            ref = m(x_all, None, None)
            torch.cuda.synchronize()
            names = ("hidden_states", "router_logits", "dynamic_top_k", "expert_mask", "global_weight")
            eq = {n: bool(torch.equal(g.reshape(r.shape), r)) for n, g, r in zip(names, gathered, ref)}
            parity["ep_equals_single_gpu"] = all(eq.values())
            parity["ep_fields_equal"] = eq
            parity["ep_max_abs_diff"] = float((gathered[0].reshape(ref[0].shape).float() - ref[0].float()).abs().max())
            parity["oracle"] = oracle_parity(m, x_all, ref, 8192)     # and the single-GPU result against the CPU oracle
            parity["routing_bit_exact"] = parity["oracle"]["routing_bit_exact"]
            parity["max_rel"] = parity["oracle"]["max_rel"]
            parity["ok"] = bool(parity["ep_equals_single_gpu"] and parity["oracle"]["ok"])
            verdict[0] = 1 if parity["ok"] else 0
            del ref
        dist.broadcast(verdict, 0)
        parity.setdefault("ok", bool(verdict[0].item()))
        del x_all, gathered
        torch.cuda.empty_cache()

    # ---- e2e: the public host-buffer API (unimoe_audio_b200.host.HostPipeline): every step copies its input
    # from pinned host memory and its whole 6-tuple back to pinned host memory inside the timed region; the
    # copies of neighbouring steps overlap the compute of the current one (3 streams, PCIe full duplex) ----
    from unimoe_audio_b200.host import HostPipeline
    x_host = [x.cpu().pin_memory() for x in xs[:2]]
    # depth 3: with 2 the H2D of step i+2 could only be submitted after the D2H of step i had finished
    pipe = HostPipeline(layer, depth=3, device=dev)
    e2e_steps = max(6, min(args.steps, 40))
    for i in range(max(args.warmup, 3) + 3):    # warm-up with the timed loop's shape (two steps in flight): the caching
        if len(pipe.pending) == pipe.depth:     # allocator then already owns every block the pipeline cycles through
            pipe.result()
        pipe.submit(x_host[i % 2])
    while pipe.pending:
        pipe.result()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for i in range(e2e_steps):
        if len(pipe.pending) == pipe.depth:
            o = pipe.result()
        pipe.submit(x_host[i % 2])
    while pipe.pending:
        o = pipe.result()                       # host tensors of the last step are complete here
    wall_ms = (time.perf_counter() - t0) * 1e3
    e.record()
    barrier()
    e2e_ms = max(s.elapsed_time(e), wall_ms)
    # ---- the same copies WITHOUT the layer: what the host side of this box sustains with `world` ranks copying at once ----
    copy_ms = pipe.copy_only_ms(x_host[0], e2e_steps)
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms, copy_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, copy_ms = t.tolist()
    e2e_value = T * world * e2e_steps / (e2e_ms * 1e-3)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    copy_gbs_rank = (h2d + d2h) * e2e_steps / (copy_ms * 1e-3) / 1e9

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = cpu_arm(1024, reps=2, warmup=1)
        arm.pop("sample_times_s")
        cpu_baseline = {"value": arm["value"], "unit": UNIT, "cores": arm["cores"], "kind": arm["kind"],
                        "sample": "1 x 1024 tokens of the same workload, bf16, all host threads, mean of 2",
                        "config1_fp32": arm["config1_fp32"], "config1_bf16": arm["config1_bf16"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "tokens_per_step_global": T * world, "parallelism": parallelism,
                       "ep_path": ep_path,
                       "l2": f"no explicit flush: inputs rotate over {n_rot} buffers ({n_rot * T * 4096 >> 20} MiB) and each step "
                             "streams 304 MB of expert weights + >1 GB of intermediates through the 126 MB L2",
                       "ffn_rows_rank0": n_rows_local, "weights": "random N(0, 0.02^2)", "top_p": 0.7},
            "clocks": clocks,
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "wall_ms_per_step": wall_ms / e2e_steps,
                    "api": "unimoe_audio_b200.host.HostPipeline(depth=3): pinned host in/out, copies overlapped across steps",
                    "copy_only": {"ms_per_step": copy_ms / e2e_steps, "gbs_per_gpu_both_directions": copy_gbs_rank,
                                  "gbs_aggregate": copy_gbs_rank * world,
                                  "note": "the step's H2D + D2H copies alone (no kernels), all ranks at once: when this is "
                                          "as long as the e2e step, e2e is bound by the host side of the box, not the GPUs"},
                    "bound": "host copies" if copy_ms > 0.85 * e2e_ms else "gpu",
                    "cpu_binding": binding},
            "gpu_launches": (KERNELS_PER_STEP if world == 1 else layer.kernels_per_step) * args.steps,
            "roofline": roofline,
            "stage_ms": stage_ms,
            "comm_stream_ms": comm_ms,
            "rank_ms_per_step": [t / args.steps for t in rank_ms],
            "rank_time_spread": (max(rank_ms) - min(rank_ms)) / max(rank_ms) if rank_ms else 0.0,
            "host_enqueue_ms_per_step": host_ms_per_step,
            "step_ms_first3": [a.elapsed_time(b) for a, b in step_ev[:3]],
            "step_ms_last3": [a.elapsed_time(b) for a, b in step_ev[-3:]],
            "cpu_baseline": cpu_baseline,
            "peaks": peaks,
        }
        if workload_id == "c4" and world > 1:
            # the same workload on ONE GPU (`--gpus 1 --workload c4`, a separate run: it cannot be timed inside a multi-rank
            # job), so that the line carries its own same-workload scaling reference next to the driver's N = 1 line
            ref_path = os.path.join(ROOT, "profiles", "r02_bench_n1_configs3_steps10.json")
            try:
                with open(ref_path) as fh:
                    ref1 = json.loads([ln for ln in fh if ln.startswith("{")][0])
                line["same_workload_n1"] = {"value": ref1["value"], "ms_per_step": ref1["ms_per_step"], "sm_mhz": ref1["clocks"]["sm_mhz"],
                                            "source": "profiles/r02_bench_n1_configs3_steps10.json (committed earlier run, not this job)",
                                            "efficiency_vs_it": value / (world * ref1["value"])}
            except Exception:  # noqa: BLE001
                pass
        print(json.dumps(line), flush=True)
    ok = bool(parity.get("ok", False))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.stderr.write("bench.py: PARITY GATE FAILED -- the measured step did not compute the reference's result\n")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["auto", "c2", "c4"], default="auto",
                    help="c2 = BASELINE.json configs[1] (8 x 2048 tokens, one GPU), c4 = configs[3] (global 64 x 4096 tokens); "
                         "auto = c2 on one GPU, c4 on several")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
