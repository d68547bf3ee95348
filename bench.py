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

The other BASELINE.json configs are measured in the same run and reported under
`extra.configs` (device-resident, CUDA events, max over ranks): cfg4 -- the
multi-GPU config, Duvenaud + Kipf stack on molecular graphs, at every N both
with the global batch of 8192 graphs sharded over the ranks (strong) and with
8192 graphs per rank (weak) -- and, at N = 1, cfg1 (chemical Duvenaud), cfg3
(2 M-vertex inference), cfg5 (power law) and a ragged variant of cfg2.
`--extras none|cfg4|all` selects them (default: all at N = 1, cfg4 at N > 1).
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
def kernel_bytes(tag, V, Z, tcg=True):
    """Bytes a launch has to move (what the kernel reads and writes, each tensor once).
    The tensor-core gather kernels read the adjacency as 16-byte bit masks per vertex
    (tcg=True); the list kernels read row_ptr + one neighbour byte per entry."""
    vf = 4 * V * F
    idx = 16 * V if tcg else 4 * (V + 1) + Z
    table = {
        # fused propagate + transform + activation: adjacency + deg^-1/2 + X read, the saved
        # aggregate P and H written, 8 B of sign bits per row
        "pipe_gather_fwd": idx + 4 * V + vf + 2 * vf + 8 * V + 4 * F * F,
        # fused last layer + MSE: X and target read, P and the loss gradient written
        "pipe_gather_fwd_mse": idx + 4 * V + 2 * vf + 2 * vf + 4 * V + 4 * F * F,
        # fused dP = gY W^T, CSC gather, .* relu'(H): gY read, gY_{t-1} written, relu' from the
        # sign bits the forward recorded (8 B per row instead of the 4VF of saved activations)
        "pipe_gather_bwd": idx + 2 * vf + 8 * V + 4 * F * F,
        # the dW products of both layers run in ONE launch (P and gY read, per layer)
        "pipe_tn": 2 * (2 * vf + 4 * F * F),
        "pipe_tn_reduce": 4 * F * F,
        "aggregate_v4_g16_c1_coef": 4 * (V + 1) + 8 * Z + 2 * vf,
        "aggregate_v4_g16_c1": 4 * (V + 1) + 4 * Z + 2 * vf,
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


def survey_kernel_bytes(tag, V, Z):
    """The SURVEY 8(d) contract figure (int32 index lists) for the same launches, reported
    beside the tighter count above."""
    vf = 4 * V * F
    idx = 4 * (V + 1) + 4 * Z
    return {"pipe_gather_fwd": idx + 4 * V + 3 * vf + 4 * F * F,
            "pipe_gather_fwd_mse": idx + 4 * V + 4 * vf + 4 * F * F,
            "pipe_gather_bwd": idx + 2 * vf + 8 * V + 4 * F * F}.get(tag)


def step_bytes(V, Z):
    """Compulsory bytes of one cfg2 train step (BASELINE.md section 3 formulas)."""
    fwd = 4 * (V + 1) + 4 * Z + 4 * V + 4 * V * F + 4 * V * F + 4 * V * F   # incl. saving P
    bwd_dense = 8 * V * F + 4 * V * F + 8 * F * F
    bwd_scatter = 4 * (V + 1) + 4 * Z + 8 * V * F
    return 2 * fwd + 2 * bwd_dense + bwd_scatter


def run_extra_configs(which, ab, L, rank, world, barrier, max_over_ranks):
    """Device-resident timings of the other BASELINE.json configs (CUDA events on the library
    stream, barrier on both sides, max over ranks).  Returns {name: {...}} (rank 0 keeps it)."""
    from athena_b200 import synth
    out = {}

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        barrier()
        ms = C.c_float()
        ab.check(L.athena_cuda_timer_start(1))
        for _ in range(steps):
            step()
        ab.check(L.athena_cuda_timer_stop(1, C.byref(ms)))
        barrier()
        return max_over_ranks(float(ms.value)) / steps

    def launches_and_kernels(step, n=3):
        n0 = np.zeros(1, np.int64); n1 = np.zeros(1, np.int64)
        ab.check(L.athena_cuda_synchronize())
        L.athena_cuda_launch_count(ab.ptr(n0))
        ab.check(L.athena_cuda_profile_begin())
        for _ in range(n):
            step()
        nt = C.c_int32()
        ab.check(L.athena_cuda_profile_end(C.byref(nt)))
        L.athena_cuda_launch_count(ab.ptr(n1))
        ks = {}
        name = C.create_string_buffer(96)
        cnt = C.c_int64(); tms = C.c_float()
        for i in range(nt.value):
            ab.check(L.athena_cuda_profile_get(i, name, 96, C.byref(cnt), C.byref(tms)))
            ks[name.value.decode()] = {"launches_per_step": cnt.value / n,
                                       "us_per_launch": round(tms.value / cnt.value * 1e3, 2)}
        return int(n1[0] - n0[0]) / n, ks

    def cfg4(mode):
        # Kipf(32->32) x 2 + Duvenaud(T=2, 6 degree buckets, 32 outputs), Adam; the Duvenaud
        # layer reads the ORIGINAL edge features (SURVEY finding 7)
        G = 8192
        if mode == "strong":     # one global batch of 8192 graphs sharded by graph
            rng = np.random.default_rng(2)
            p_all = synth.molecular_batch(G, 32, 4, rng)
            tgt_all = rng.random((G, 32)).astype(np.float32)
            first = np.zeros(world + 1, np.int32)
            nz64 = p_all.nz.astype(np.int64)
            ab.check(L.athena_cuda_shard_graphs(p_all.B, ab.ptr(nz64), world, ab.ptr(first)))
            g0, g1 = int(first[rank]), int(first[rank + 1])
            q, tgt, global_B = p_all.slice(g0, g1), np.ascontiguousarray(tgt_all[g0:g1]), G
            Zg, Vg = p_all.Z, p_all.V
        else:                    # 8192 graphs per rank
            rng = np.random.default_rng(200 + rank)
            q = synth.molecular_batch(G, 32, 4, rng)
            tgt = rng.random((G, 32)).astype(np.float32)
            global_B = G * world
            Zg = Vg = None           # summed over the ranks below
        net = ab.network_type()
        net.add(ab.kipf_msgpass_layer_type([32, 32], 1, "relu"))
        net.add(ab.kipf_msgpass_layer_type([32, 32], 1, "relu"))
        net.add(ab.duvenaud_msgpass_layer_type([32], [4], 2, 6, 32))
        net.compile(ab.adam_optimiser_type(0.001), batch_size=q.B)
        net.set_params((np.random.default_rng(7).standard_normal(net.num_params) * 0.1)
                       .astype(np.float32))
        batch = ab.GraphBatch(q)
        x = ab.DeviceArray.from_host(q.x)
        e = ab.DeviceArray.from_host(q.e)
        t = ab.DeviceArray.from_host(tgt)

        def step():
            ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x), ab.ptr(e),
                                                      ab.ptr(t), ab.MEM_DEVICE, global_B, None))
        ms = timed(step, 30, 5)
        lps, ks = launches_and_kernels(step)
        if mode != "strong":
            Zg, Vg = int(_sum_over_ranks(q.Z)), int(_sum_over_ranks(q.V))
        loss = C.c_float()
        ab.check(L.athena_cuda_network_last_loss(net.handle, C.byref(loss)))
        return {"workload": "cfg4: Kipf(32->32, relu) x 2 + Duvenaud(T=2, D=6, 32 outputs, sigmoid / "
                            "softmax) train, molecular graphs (V~U[10,50], degree<=4+self, Fv=32, "
                            "Fe=4), MSE, Adam; fwd+bwd+exchange+step",
                "scaling": mode, "global_graphs": global_B, "graphs_this_rank": q.B,
                "global_vertices": Vg, "global_entries": Zg, "ms_per_step": ms,
                "graphs_per_s": global_B / ms * 1e3, "edges_per_s": Zg / ms * 1e3,
                "launches_per_step": lps, "kernels": ks, "final_loss": float(loss.value)}

    def cfg1():
        rng = np.random.default_rng(42)
        q = synth.chemical_batch(8, rng)
        net = ab.network_type()
        net.add(ab.duvenaud_msgpass_layer_type([6], [1], 4, 10, 10))
        net.compile(ab.adam_optimiser_type(0.01, clip_dict=ab.clip_type(clip_norm=0.1)),
                    batch_size=8)
        net.set_params((rng.standard_normal(net.num_params) * 0.3).astype(np.float32))
        batch = ab.GraphBatch(q)
        x = ab.DeviceArray.from_host(q.x)
        e = ab.DeviceArray.from_host(q.e)
        t = ab.DeviceArray.from_host(rng.random((q.B, 10)).astype(np.float32))

        def step():
            ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x), ab.ptr(e),
                                                      ab.ptr(t), ab.MEM_DEVICE, q.B, None))
        ms = timed(step, 200, 20)
        lps, ks = launches_and_kernels(step, 10)
        return {"workload": "cfg1: example/msgpass_chemical dims -- Duvenaud (T=4, 10 degree "
                            "buckets, 10 outputs), 8 graphs x 8 atoms per batch (synthetic "
                            "stand-in), Adam + clip_norm 0.1; one train step",
                "vertices": q.V, "entries": q.Z, "us_per_step": ms * 1e3,
                "graphs_per_s": q.B / ms * 1e3, "edges_per_s": q.Z / ms * 1e3,
                "launches_per_step": lps, "kernels": ks, "roofline": "n/a (latency-bound)"}

    def kipf2(q, Fw, what, train=True, steps=20):
        rng = np.random.default_rng(5)
        net = ab.network_type()
        net.add(ab.kipf_msgpass_layer_type([Fw, Fw], 1, "relu"))
        net.add(ab.kipf_msgpass_layer_type([Fw, Fw], 1, "none"))
        net.compile(ab.sgd_optimiser_type(LR), batch_size=q.B)
        net.set_params((rng.standard_normal(net.num_params) / np.sqrt(Fw)).astype(np.float32))
        batch = ab.GraphBatch(q)
        x = ab.DeviceArray.from_host(q.x)
        if train:
            t = ab.DeviceArray.from_host(rng.standard_normal((q.V, Fw)).astype(np.float32))

            def step():
                ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x), None,
                                                          ab.ptr(t), ab.MEM_DEVICE, q.B, None))
        else:
            o = ab.DeviceArray((q.V, Fw))

            def step():
                ab.check(L.athena_cuda_network_forward(net.handle, batch.handle, ab.ptr(x), None,
                                                       ab.ptr(o), ab.MEM_DEVICE))
        ms = timed(step, steps, 5)
        lps, ks = launches_and_kernels(step)
        return {"workload": what, "vertices": q.V, "entries": q.Z, "graphs": q.B,
                "ms_per_step": ms, "edges_per_s": q.Z / ms * 1e3, "launches_per_step": lps,
                "kernels": ks}

    for name in which:
        try:
            if name == "cfg4":
                out["cfg4"] = cfg4("strong")
                out["cfg4_weak"] = cfg4("weak")
            elif name == "cfg1":
                out["cfg1"] = cfg1()
            elif name == "cfg2_ragged":
                q = synth.ragged_batch(GRAPHS, 33, 64, HALF_DEG, F, np.random.default_rng(11))
                out[name] = kipf2(q, F, "cfg2 with ragged graphs: Kipf 2 x (64->64) train, "
                                  f"{GRAPHS} graphs of 33..64 vertices, 12 neighbours + self loop "
                                  "(graph-aligned tiles partially filled, unequal)")
            elif name == "cfg3":
                q = synth.random_graph(2_000_000, 8, 128, np.random.default_rng(1))
                r = kipf2(q, 128, "cfg3: Kipf 2 x (128->128) inference, one graph, V = 2e6, "
                          "16 neighbours + self loop", train=False)
                V, Z = q.V, q.Z
                comp = 2 * (4 * (V + 1) + 4 * Z + 4 * V + 8 * V * 128)
                gath = 2 * (Z * (4 + 4 * 128) + V * (8 + 4 * 128))
                peak, _ = peaks()
                r["roofline"] = {"bound": "hbm", "compulsory_bytes": comp, "gather_model_bytes": gath,
                                 "compulsory_frac": comp / (r["ms_per_step"] * 1e-3) / 1e9 / peak,
                                 "gather_model_frac": gath / (r["ms_per_step"] * 1e-3) / 1e9 / peak}
                out[name] = r
            elif name == "cfg5":
                q = synth.powerlaw_batch(64, 16384, 64, np.random.default_rng(3), max_degree=10000)
                out[name] = kipf2(q, 64, "cfg5: Kipf 2 x (64->64) train, 64 power-law graphs x "
                                  "16384 vertices (Zipf 2.1, max degree 10000)")
        except Exception as exc:  # noqa: BLE001  (an extra must never take the headline down)
            out[name] = {"error": str(exc)[:300]}
    return out


_sum_over_ranks = None


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
              f"({p.Z} CSR entries), full train step; C restatement of the reference at -O3 on "
              "one core (the reference is single-threaded)")
    print(json.dumps({
        "impl": "reference", "metric": "msgpass train edges/sec (fwd+bwd)", "value": value,
        "unit": "edges/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: Kipf GCN 2-layer (64->64 relu, 64->64), "
                               f"{GRAPHS} graphs x {NV} vertices per GPU, 12 neighbours + "
                               "self loop, F=64, MSE, SGD; fwd+bwd+allreduce+step",
                   "graphs_per_gpu": sample_graphs, "vertices_per_gpu": p.V,
                   "entries_per_gpu": p.Z, "sample": sample},
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
    ap.add_argument("--ref-graphs", type=int, default=GRAPHS)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--extras", default="auto", choices=["auto", "none", "cfg4", "all"])
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
                "survey_formula_bytes_per_launch": survey_kernel_bytes(top, V, Z),
                "us_per_launch": kernels[top]["us_per_launch"],
                "step": {"algorithmic_bytes": step_bytes(V, Z),
                         "achieved": step_bytes(V, Z) / (ms_step * 1e-3) / 1e9,
                         "frac": step_bytes(V, Z) / (ms_step * 1e-3) / 1e9 / peak}}
    clocks = sampler.stop() if rank == 0 else None
    barrier()

    # ---- leg 2: end to end through the public API with HOST buffers ------------------
    e2e = None
    if not args.no_e2e:
        from athena_b200 import synth
        nbuf = 2
        host = []
        for i in range(nbuf):
            q, tg = (p, target_h) if i == 0 else make_workload(rank + 100 * i)
            ne_e, il = synth.edge_lists(q)
            h = {"nv": q.nv, "ne": q.ne, "nz": q.nz, "ne_e": ne_e}
            for key, arr in (("ia", q.ia), ("ja", q.ja), ("x", q.x), ("t", tg), ("il", il)):
                buf = ab.pinned_empty(arr.shape, arr.dtype)
                buf[...] = arr
                h[key] = buf
            host.append(h)
        lossf = C.c_float()

        # (a) the reference's own flow: the reader hands over EDGE LISTS, generate_adjacency +
        #     add_self_loops build the CSR -- here on the device (athena_cuda_batch_create_from_edges)
        def e2e_step_edges(i):
            h = host[i % nbuf]
            bh = C.c_int64()
            ab.check(L.athena_cuda_batch_create_from_edges(
                C.byref(bh), B, ab.ptr(h["nv"]), ab.ptr(h["ne_e"]), ab.ptr(h["il"]),
                ab.ptr(h["nz"]), 1, ab.MEM_HOST, 0))
            ab.check(L.athena_cuda_network_train_step(net.handle, bh.value, ab.ptr(h["x"]), None,
                                                      ab.ptr(h["t"]), ab.MEM_HOST, global_B,
                                                      C.byref(lossf)))
            ab.check(L.athena_cuda_batch_destroy(bh.value))

        # (b) graphs that arrive with their CSR already built on the host (adj_ia / adj_ja)
        def e2e_step_csr(i):
            h = host[i % nbuf]
            bh = C.c_int64()
            ab.check(L.athena_cuda_batch_create(C.byref(bh), B, ab.ptr(h["nv"]), ab.ptr(h["ne"]),
                                                ab.ptr(h["nz"]), ab.ptr(h["ia"]), ab.ptr(h["ja"]),
                                                ab.MEM_HOST, 0))
            ab.check(L.athena_cuda_network_train_step(net.handle, bh.value, ab.ptr(h["x"]), None,
                                                      ab.ptr(h["t"]), ab.MEM_HOST, global_B,
                                                      C.byref(lossf)))
            ab.check(L.athena_cuda_batch_destroy(bh.value))

        def time_e2e(step_fn, keys):
            for i in range(max(3, args.warmup // 2)):
                step_fn(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                step_fn(i)
            ab.check(L.athena_cuda_synchronize())
            dt = max_over_ranks(time.perf_counter() - t0)
            barrier()
            h2d = sum(host[0][k].nbytes for k in keys) + 4 * 4 * (B + 1)
            return {"value": world * Z * args.steps / dt, "unit": "edges/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "ms_per_step": dt / args.steps * 1e3,
                    "host_to_device_GBps_all_ranks": world * h2d / (dt / args.steps) / 1e9}

        e_edges = time_e2e(e2e_step_edges, ("il", "x", "t"))
        e_csr = time_e2e(e2e_step_csr, ("ia", "ja", "x", "t"))
        # headline = the faster of the two public routes (both ship every byte of the step from
        # pinned host memory): graph_type samples with the CSR built by the caller (adj_ia /
        # adj_ja, as set_graph receives them) or EDGE LISTS with generate_adjacency +
        # add_self_loops done on the device (10 % fewer bytes; the per-graph entry counts are
        # passed as a hint, as a training loop knows them from its first epoch, so the build
        # stays asynchronous)
        e_csr["input"] = ("adj_ia / adj_ja + features + targets from pinned host memory "
                          "(athena_cuda_batch_create + athena_cuda_network_train_step)")
        e_edges["input"] = ("edge lists (index_list) + features + targets from pinned host memory; "
                            "CSR built on the device (athena_cuda_batch_create_from_edges + "
                            "athena_cuda_network_train_step)")
        best, other = (e_csr, e_edges) if e_csr["value"] >= e_edges["value"] else (e_edges, e_csr)
        e2e = dict(best)
        e2e["other_route"] = other

        # the platform's own ceiling for this step: every rank copies the same pinned buffers to
        # its device with nothing else going on (all ranks at once) -- what the host's memory
        # system and PCIe topology deliver to N GPUs, however the step is organised
        def h2d_only():
            for key, dst in (("x", x_d), ("t", t_d)):
                ab.check(L.athena_cuda_memcpy_h2d(C.c_void_p(dst.addr), ab.ptr(host[0][key]),
                                                  host[0][key].nbytes))
        for _ in range(3):
            h2d_only()
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            h2d_only()
        ab.check(L.athena_cuda_synchronize())
        dt_copy = max_over_ranks(time.perf_counter() - t0)
        barrier()
        copy_bytes = host[0]["x"].nbytes + host[0]["t"].nbytes
        e2e["host_ceiling"] = {
            "h2d_GBps_all_ranks": world * copy_bytes * 10 / dt_copy / 1e9,
            "ms_per_step_at_ceiling": e2e["h2d_bytes_per_step"] / (copy_bytes * 10 / dt_copy) * 1e3,
            "how": "every rank copies its pinned feature + target buffers (134 MB) to its device "
                   "10 times, all ranks at once, nothing else running"}

        # (c) dataset-resident epochs (network%train is handed the whole data set once,
        #     athena_network_sub.f90:3564-3565): batches and features stay on the device, a
        #     step moves nothing but the loss.  Reported beside, never instead of, the above.
        t0 = time.perf_counter()
        for i in range(args.steps):
            ab.check(L.athena_cuda_network_train_step(net.handle, batch.handle, ab.ptr(x_d), None,
                                                      ab.ptr(t_d), ab.MEM_DEVICE, global_B,
                                                      C.byref(lossf)))
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e["dataset_resident"] = {"value": world * Z * args.steps / dt, "unit": "edges/s",
                                   "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                                   "ms_per_step": dt / args.steps * 1e3}

    # ---- the other BASELINE.json configs ----------------------------------------------
    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    global _sum_over_ranks
    _sum_over_ranks = sum_over_ranks
    sel = args.extras
    if sel == "auto":
        sel = "all" if world == 1 else "cfg4"
    which = {"none": [], "cfg4": ["cfg4"],
             "all": ["cfg4", "cfg1", "cfg2_ragged", "cfg5", "cfg3"]}[sel]
    extra_cfgs = run_extra_configs(which, ab, L, rank, world, barrier, max_over_ranks) if which else {}

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
            "extra": {"configs": extra_cfgs},
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
