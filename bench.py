#!/usr/bin/env python
"""Benchmark of the message-passing hot path (BASELINE.json metric:
"msgpass train edges/sec (fwd+bwd)", workload configs[1]: Kipf GCN 2-layer,
4096 synthetic graphs x 64 nodes, avg degree 12, 64 features).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

One process per GPU (torchrun for N > 1).  A step = forward + backward
(+ gradient all-reduce) + optimiser step over one mini-batch (per-rank batch is
fixed: weak scaling).  `value` is timed with CUDA events on the library stream
with inputs resident in HBM; `e2e` goes through the public API with pinned HOST
buffers (batch build + H2D copies + loss read-back inside the timed region).
An edge is one CSR entry, self-loops included.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F = 64
GRAPHS, NV, HALF_DEG = 4096, 64, 6
LR = 0.01


def env_int(k, d):
    return int(os.environ.get(k, d))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(rank, graphs=GRAPHS):
    from athena_b200 import synth
    rng = np.random.default_rng(1000 + rank)
    p = synth.regular_batch(graphs, NV, HALF_DEG, F, rng)
    target = rng.standard_normal((p.V, F), dtype=np.float32)
    return p, target


def init_params(n):
    rng = np.random.default_rng(7)
    return (rng.standard_normal(n) / np.sqrt(F)).astype(np.float32)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line).
    Sampled through NVML every 2 ms (the timed region of the default run is only tens of
    milliseconds, too short for `nvidia-smi -lms`); falls back to nvidia-smi when the NVML
    binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.sm, self.reasons, self.power = [], [], []   # (time, value) samples
        self.mx = None
        self.window = [None, None]
        self._stop = threading.Event()
        self.t = None
        self.how = None

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                "sw_power_cap": 0x4}
        k = 0
        while not self._stop.is_set():
            try:
                t = time.perf_counter()
                self.sm.append((t, float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.reasons.append((t, [name for name, bit in bits.items() if r & bit]))
                if k % 8 == 0:
                    self.power.append((t, nv.nvmlDeviceGetPowerUsage(h) / 1e3))
                k += 1
            except Exception:
                pass
            time.sleep(0.001)

    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            t = time.perf_counter()
            try:
                self.sm.append((t, float(f[1])))
                self.mx = float(f[2])
                self.power.append((t, float(f[3])))
            except ValueError:
                continue
            self.reasons.append((t, [name for name, val in zip(names, f[5:9])
                                     if val.lower().startswith("active")]))

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.device
            if vis:
                try:
                    idx = int(vis.split(",")[self.device])
                except Exception:
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.how = "nvml, ~1 ms period"
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi -lms 20"
            self.t = threading.Thread(target=self._smi_loop, daemon=True)
            self.t.start()
        except Exception:
            self.how = None

    def mark_begin(self):
        self.window[0] = time.perf_counter()

    def mark_end(self):
        self.window[1] = time.perf_counter()

    def stop(self):
        self._stop.set()
        if getattr(self, "proc", None):
            self.proc.terminate()
        if self.t:
            self.t.join(timeout=2)
        t0, t1 = self.window
        inside = lambda t: (t0 is None or t >= t0) and (t1 is None or t <= t1)
        sm = [v for t, v in self.sm if inside(t)]
        where = "timed region"
        if len(sm) < 3:   # region shorter than a few sampling periods: widen to the loaded span
            sm = [v for t, v in self.sm if t0 is None or t >= t0 - 0.05]
            where = "timed region and the identical per-kernel timing pass right after it"
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "samples": 0,
                    "reasons": ["no clock sample (nvml and nvidia-smi unavailable)"]}
        reasons = sorted({r for t, rs in self.reasons if t0 is None or t >= t0 - 0.05 for r in rs})
        power = [v for t, v in self.power if t0 is None or t >= t0 - 0.05]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)),
                "sm_max_mhz": self.mx, "samples": len(sm),
                "power_w_max": max(power) if power else None, "reasons": reasons,
                "how": self.how, "where": where}


# algorithmic (compulsory) bytes per launch, SURVEY 8(d) / DESIGN.md "Roofline accounting":
# every tensor a kernel must touch counted once, int32 indices, fp32 values.
def kernel_bytes(tag, V, Z):
    vf = 4 * V * F
    idx = 4 * (V + 1) + 4 * Z
    table = {
        # fused propagate + transform + activation (SURVEY B_f with the saved aggregate P):
        # row_ptr + col + degree/coefficient vector + X read, P and H written
        "pipe_gather_fwd": idx + 4 * V + vf + 2 * vf + 4 * F * F,
        # fused last layer + MSE: X and target read, P and the loss gradient written
        "pipe_gather_fwd_mse": idx + 4 * V + 2 * vf + 2 * vf + 4 * F * F,
        # fused dP = gY W^T, CSC gather, .* relu'(H): gY read, gY_{t-1} written, relu' from the
        # sign bits the forward recorded (8 B per row instead of the 4VF of saved activations)
        "pipe_gather_bwd": idx + 2 * vf + 8 * V + 4 * F * F,
        # the dW products of both layers run in ONE launch (P and gY read, per layer)
        "pipe_tn": 2 * (2 * vf + 4 * F * F),
        "pipe_tn_reduce": 4 * F * F,
        "aggregate_v4_g16_c1_coef": idx + 4 * V + 2 * vf,
        "aggregate_v4_g16_c1": idx + 2 * vf,
        "gemm_nn": 2 * vf + 4 * F * F,
        "gemm_nt": 2 * vf + 4 * F * F,
        "gemm_tn_partial": 2 * vf,
        "gemm_tn_reduce": 4 * F * F,
        "act_bwd": 3 * vf,
        "mse_graph": 3 * vf,
    }
    return table.get(tag)


def measured_traffic(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `tag` from the committed
    `ncu --set full` capture (profiles/traffic.json, written by tools/ncu_summary.py)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(tag, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def step_bytes(V, Z):
    """Compulsory bytes of one cfg2 train step (BASELINE.md section 3 formulas)."""
    fwd = 4 * (V + 1) + 4 * Z + 4 * V + 4 * V * F + 4 * V * F + 4 * V * F   # incl. saving P
    bwd_dense = 8 * V * F + 4 * V * F + 8 * F * F
    bwd_scatter = 4 * (V + 1) + 4 * Z + 8 * V * F
    return 2 * fwd + 2 * bwd_dense + bwd_scatter


def run_reference(args, rank, world):
    """The reference's CPU path for the same step.  The Fortran build is impossible in this
    image (no Fortran compiler, un-vendored deps), so this times the line-by-line C
    restatement (oracle/, -O3) on ONE core: the reference has no working threading
    (SURVEY 0.1), so one core is all the host threads it can use."""
    if rank != 0:
        return
    from oracle.oracle import Batch, LayerSpec, OptimSpec, Oracle
    o = Oracle("fast")
    sample_graphs = args.ref_graphs
    p, target = make_workload(0, sample_graphs)
    specs = [LayerSpec("kipf", [F, F], 1, activation="relu"),
             LayerSpec("kipf", [F, F], 1, activation="none")]
    n = o.num_params(specs)
    params = init_params(n)
    s1 = np.zeros(n, np.float32); s2 = np.zeros(n, np.float32)
    b = Batch(p.nv, p.ne, p.ia, p.ja, p.x, None)
    opt = OptimSpec("sgd", lr=LR)
    it = 0
    for _ in range(args.warmup):
        it += 1
        o.train_step(specs, params, b, target, opt, s1, s2, it)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        it += 1
        o.train_step(specs, params, b, target, opt, s1, s2, it)
    dt = time.perf_counter() - t0
    value = p.Z * args.steps / dt
    sample = (f"{sample_graphs} of the {GRAPHS} graphs of the cfg2 batch per step "
              f"({p.Z} CSR entries), full train step")
    print(json.dumps({
        "impl": "reference", "metric": "msgpass train edges/sec (fwd+bwd)", "value": value,
        "unit": "edges/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: Kipf GCN 2-layer (64->64 relu, 64->64), 64-vertex graphs, "
                               "12 neighbours + self loop, F=64, MSE, SGD", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": 1, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-graphs", type=int, default=128)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # Bind the process to the CPUs (NUMA node) closest to its GPU before any pinned host
    # buffer is allocated: with several ranks on one box the host->device copies of the
    # end-to-end leg otherwise cross the socket interconnect.
    if world > 1:
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local]) if vis else local
            nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(idx))
        except Exception:
            pass

    import athena_b200 as ab
    L = ab.lib()
    ab.check(L.athena_cuda_init(local))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idbuf = [None]
        if rank == 0:
            raw = C.create_string_buffer(ab._lib.COMM_ID_BYTES)
            ab.check(L.athena_cuda_comm_unique_id(raw))
            idbuf = [raw.raw]
        dist.broadcast_object_list(idbuf, src=0)
        ab.check(L.athena_cuda_comm_init(world, rank, C.create_string_buffer(idbuf[0], 128)))
        use_p2p = os.environ.get("ATHENA_BENCH_NO_P2P", "0") == "0" and world <= 8
        if use_p2p:
            # peer-memory gradient exchange fused with the step.  Every rank must agree: if the
            # IPC set-up fails anywhere, all ranks fall back to the NCCL all-reduce.
            ok = 1
            try:
                mine = C.create_string_buffer(ab._lib.P2P_HANDLE_BYTES)
                ab.check(L.athena_cuda_comm_p2p_export(mine))
                allh = [None] * world
                dist.all_gather_object(allh, mine.raw)
                ab.check(L.athena_cuda_comm_p2p_import(world, rank,
                                                       C.create_string_buffer(b"".join(allh))))
            except Exception as exc:  # noqa: BLE001
                print(f"[rank {rank}] peer-memory exchange unavailable: {exc}", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                use_p2p = False
                ab.check(L.athena_cuda_comm_destroy())
                idbuf = [None]
                if rank == 0:
                    raw = C.create_string_buffer(ab._lib.COMM_ID_BYTES)
                    ab.check(L.athena_cuda_comm_unique_id(raw))
                    idbuf = [raw.raw]
                dist.broadcast_object_list(idbuf, src=0)
                ab.check(L.athena_cuda_comm_init(world, rank, C.create_string_buffer(idbuf[0], 128)))
        exchange = "peer-memory sum fused with the step, NVLink" if use_p2p else "NCCL all-reduce"
    else:
        exchange = "none"

    def barrier():
        ab.check(L.athena_cuda_synchronize())
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload ----------------------------------------------------------------
    p, target_h = make_workload(rank)
    V, Z, B = p.V, p.Z, p.B
    global_B = B * world
    net = ab.network_type()
    net.add(ab.kipf_msgpass_layer_type([F, F], 1, "relu"))
    net.add(ab.kipf_msgpass_layer_type([F, F], 1, "none"))
    net.compile(ab.sgd_optimiser_type(LR), batch_size=B)
    net.set_params(init_params(net.num_params))

    # ---- leg 1: inputs resident in HBM -----------------------------------------------
    x_d = ab.DeviceArray.from_host(p.x)
    t_d = ab.DeviceArray.from_host(target_h)
    batch = ab.GraphBatch(p)

    def dev_step():
        ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x_d), None,
                                                  ab.ptr(t_d), ab.MEM_DEVICE, global_B, None))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        dev_step()
    barrier()
    n0 = np.zeros(1, np.int64); n1 = np.zeros(1, np.int64)
    L.athena_cuda_launch_count(ab.ptr(n0))
    ms = C.c_float()
    sampler.mark_begin()
    ab.check(L.athena_cuda_timer_start(0))
    for _ in range(args.steps):
        dev_step()
    ab.check(L.athena_cuda_timer_stop(0, C.byref(ms)))
    sampler.mark_end()
    barrier()
    L.athena_cuda_launch_count(ab.ptr(n1))
    launches = int(n1[0] - n0[0])
    ms_total = max_over_ranks(float(ms.value))
    ms_step = ms_total / args.steps
    value = world * Z * args.steps / (ms_total / 1e3)
    loss = C.c_float()
    ab.check(L.athena_cuda_network_last_loss(net.handle, C.byref(loss)))

    # ---- per-kernel durations (same steps, events after every launch) -------------
    roof = None
    kernels = {}
    # (every rank takes the same steps -- each one contains the gradient all-reduce -- and
    # rank 0 reports its own per-kernel times)
    prof_steps = max(3, min(args.steps, 10))
    ab.check(L.athena_cuda_synchronize())
    ab.check(L.athena_cuda_profile_begin())
    for _ in range(prof_steps):
        dev_step()
    ntags = C.c_int32()
    ab.check(L.athena_cuda_profile_end(C.byref(ntags)))
    if rank == 0:
        name = C.create_string_buffer(96)
        cnt = C.c_int64(); tms = C.c_float()
        for i in range(ntags.value):
            ab.check(L.athena_cuda_profile_get(i, name, 96, C.byref(cnt), C.byref(tms)))
            kernels[name.value.decode()] = {"launches_per_step": cnt.value / prof_steps,
                                            "us_per_launch": tms.value / cnt.value * 1e3,
                                            "ms_per_step": tms.value / prof_steps}
        peak, peak_src = peaks()
        tot = sum(k["ms_per_step"] for k in kernels.values())
        for k, v in kernels.items():
            v["share"] = v["ms_per_step"] / tot
            nb = kernel_bytes(k, V, Z)
            if nb:
                v["algorithmic_GBps"] = nb / (v["us_per_launch"] * 1e-6) / 1e9
        top = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        nb = kernel_bytes(top, V, Z)
        ach = nb / (kernels[top]["us_per_launch"] * 1e-6) / 1e9 if nb else None
        roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None, "traffic": measured_traffic(top),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": nb,
                "us_per_launch": kernels[top]["us_per_launch"],
                "step": {"algorithmic_bytes": step_bytes(V, Z),
                         "achieved": step_bytes(V, Z) / (ms_step * 1e-3) / 1e9,
                         "frac": step_bytes(V, Z) / (ms_step * 1e-3) / 1e9 / peak}}
    clocks = sampler.stop() if rank == 0 else None
    barrier()

    # ---- leg 2: end to end through the public API with HOST buffers ------------------
    e2e = None
    if not args.no_e2e:
        nbuf = 2
        host = []
        for i in range(nbuf):
            q, tg = (p, target_h) if i == 0 else make_workload(rank + 100 * i)
            h = {"nv": q.nv, "ne": q.ne, "nz": q.nz}
            for key, arr in (("ia", q.ia), ("ja", q.ja), ("x", q.x), ("t", tg)):
                buf = ab.pinned_empty(arr.shape, arr.dtype)
                buf[...] = arr
                h[key] = buf
            host.append(h)
        h2d = sum(host[0][k].nbytes for k in ("ia", "ja", "x", "t")) + 5 * 4 * (B + 1)
        lossf = C.c_float()

        def e2e_step(i):
            h = host[i % nbuf]
            bh = C.c_int64()
            ab.check(L.athena_cuda_batch_create(C.byref(bh), B, ab.ptr(h["nv"]), ab.ptr(h["ne"]),
                                                ab.ptr(h["nz"]), ab.ptr(h["ia"]), ab.ptr(h["ja"]),
                                                ab.MEM_HOST, 0))
            ab.check(L.athena_cuda_network_train_step(net.handle, bh.value, ab.ptr(h["x"]), None,
                                                      ab.ptr(h["t"]), ab.MEM_HOST, global_B,
                                                      C.byref(lossf)))
            ab.check(L.athena_cuda_batch_destroy(bh.value))

        for i in range(max(3, args.warmup // 2)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        ab.check(L.athena_cuda_synchronize())
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e = {"value": world * Z * args.steps / dt, "unit": "edges/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
               "ms_per_step": dt / args.steps * 1e3}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.oracle import Batch, LayerSpec, OptimSpec, Oracle
        o = Oracle("fast")
        specs = [LayerSpec("kipf", [F, F], 1, activation="relu"),
                 LayerSpec("kipf", [F, F], 1, activation="none")]
        g = 64
        sp = p.slice(0, g)
        bt = Batch(sp.nv, sp.ne, sp.ia, sp.ja, sp.x, None)
        nparam = o.num_params(specs)
        prm = init_params(nparam)
        s1 = np.zeros(nparam, np.float32); s2 = np.zeros(nparam, np.float32)
        tg = target_h[:sp.V]
        o.train_step(specs, prm, bt, tg, OptimSpec("sgd", lr=LR), s1, s2, 1)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < args.cpu_seconds:
            reps += 1
            o.train_step(specs, prm, bt, tg, OptimSpec("sgd", lr=LR), s1, s2, 1 + reps)
        dt = time.perf_counter() - t0
        cpu = {"value": sp.Z * reps / dt, "unit": "edges/s", "cores": 1, "kind": "port",
               "host_cores_available": os.cpu_count(),
               "sample": f"first {g} graphs of the batch ({sp.Z} CSR entries) x {reps} full train "
                         f"steps, {dt:.1f} s, C restatement of the reference at -O3, one core "
                         "(the reference is single-threaded)"}

    if rank == 0:
        out = {
            "metric": "msgpass train edges/sec (fwd+bwd)", "value": value, "unit": "edges/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "cfg2: Kipf GCN 2-layer (64->64 relu, 64->64), "
                                   f"{GRAPHS} graphs x {NV} vertices per GPU, 12 neighbours + "
                                   "self loop, F=64, MSE, SGD; fwd+bwd+allreduce+step",
                       "graphs_per_gpu": B, "vertices_per_gpu": V, "entries_per_gpu": Z,
                       "parallelism": f"dp{world} (graph-sharded; gradient exchange of "
                                      f"{net.num_params + 1} floats: " +
                                      exchange + ")",
                       "l2": "working set ~1 GB per step > 126 MB L2 (no flush needed)"},
            "e2e": e2e, "gpu_launches": launches, "launches_per_step": launches / args.steps,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
            "final_loss": float(loss.value),
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
