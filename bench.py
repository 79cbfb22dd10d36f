"""Benchmark of the circuit-execution hot path on B200 (BASELINE.json metric:
circuit evals/sec, batched).

Workload (BASELINE.json configs[1], SURVEY.md section 8(d) cfg 2, grid B "dense"):
``Coefficients.get_spectrum`` on ``Model(n_qubits=4, n_layers=4, "Hardware_Efficient")``,
expval on all 4 qubits, 1024 parameter samples (NumPy default_rng(1000), uniform[0, 2pi)) x a
264-point input grid (``mfs=8``) = 270 336 circuit evaluations per step, complex128.

A STEP (identical at every GPU count) = the circuit kernels (k_pre x2, k_reg) over the batch
resident in HBM, the grid DFT (k_grid_dft: mean over qubits + transform, what get_spectrum
returns) and the FCC sufficient statistics over the samples (k_coef_moments); with N > 1 the
statistics (a few KB) cross ranks in a one-shot all-reduce over peer-mapped symmetric memory
(k_allreduce_oneshot) - no library collective inside the step.

One JSON line on stdout (rank 0):
  value     evals/s of the step, arguments resident in HBM (CUDA events, graph replay)
  e2e       evals/s through Coefficients.get_spectrum(model, mfs=8, ...) with host NumPy
            params / grid in and the host coefficient array out (H2D + D2H in the timed region)
  roofline  dominant kernel (k_reg<double,4>) against the FP64 FMA peak measured on the same
            GPU by qmlb_fma_peak (MEASURED_PEAKS.json has no FMA figure)
  cpu_baseline  the reference-faithful CPU restatement (oracle: one batched einsum per tape
            op, torch-CPU complex128, all host threads, + numpy FFT) on the same workload
  config_1 / config_3 / config_4   device-resident evals/s of the other small-n BASELINE
            configs (k_reg, k_frame single-CTA, k_frame cluster)
  gate_pass config 5 at n = 32 (k_fstream tile passes): HBM GB/s per pass, s per circuit
``--impl reference`` times only the CPU restatement (the reference itself needs JAX,
which this image does not have - DESIGN.md).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = ("BASELINE configs[1] / SURVEY 8(d) cfg2 grid B: Coefficients.get_spectrum on "
            "Model(4,4,'Hardware_Efficient'), expval on 4 qubits, 1024 param samples "
            "(default_rng(1000+rank)) x 264-point input grid per GPU")
N_QUBITS, N_LAYERS, ANSATZ = 4, 4, "Hardware_Efficient"
N_PARAM_SAMPLES, MFS = 1024, 8
ALGO_FLOP_PER_EVAL = 26624  # 76 one-qubit + 20 CX dense gates, n = 4
L2_FLUSH_BYTES = 512 * 1024 * 1024


def workload(seed: int = 1000):
    from qml_essentials_b200.model import Model

    model = Model(n_qubits=N_QUBITS, n_layers=N_LAYERS, circuit_type=ANSATZ)
    rng = np.random.default_rng(seed)
    params = rng.uniform(0.0, 2 * np.pi, (N_PARAM_SAMPLES, *model._params_shape))
    n_freqs = MFS * model.degree[0]
    inputs = np.arange(0.0, 2 * np.pi, 2 * np.pi / n_freqs).reshape(-1, 1)
    return model, params, inputs


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._halt = gpu_index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                     "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6
                          for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def oracle_cpu_evals_per_s(params, inputs, n_param_sample, threads, repeats=2):
    """Reference-faithful CPU restatement: all grid points x `n_param_sample` parameter
    sets, one batched einsum per tape op (torch-CPU), <Z_q> and the grid FFT of
    Coefficients._fourier_transform (numpy) - the whole get_spectrum workload."""
    from oracle import circuits as oc
    from oracle import sim as osim

    B_I, B_P = inputs.shape[0], n_param_sample
    ii, pp, _ = oc.assimilate_index(B_I, B_P)
    p_b = np.moveaxis(params[:B_P][pp], 0, -1)  # (L', P, B)
    x_b = [inputs[ii, 0]]
    tape = oc.variational_tape(N_QUBITS, N_LAYERS, ANSATZ, p_b, x_b)
    B = B_I * B_P
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        st = osim.simulate_batched(tape, N_QUBITS, B, backend="torch", threads=threads)
        probs = np.abs(st) ** 2
        pt = probs.reshape((B,) + (2,) * N_QUBITS)
        ev = np.stack([pt.sum(axis=tuple(a + 1 for a in range(N_QUBITS) if a != q))
                       @ np.array([1.0, -1.0]) for q in range(N_QUBITS)], axis=1)
        np.fft.fft(ev.reshape(B_I, B_P, N_QUBITS).mean(axis=2), axis=0)  # coefficients.py:135
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return B / best, ev.reshape(B_I, B_P, N_QUBITS), f"{B_I} grid points x {B_P} param sets"


def gate_pass_leg(ex, n_qubits, reps, dev):
    """BASELINE configs[4] / SURVEY 8(d) cfg 5: Model(n, 8, 'Hardware_Efficient'), complex64,
    one circuit, expval on all qubits; the state lives in HBM.  Timed twice (CUDA events
    around qmlb_run, state larger than L2): the tile passes of the streaming frame engine
    (k_fstream - what the product runs: every step that fits 13 bits per read+write of the
    state) and, for reference, round 1's register-group passes (k_stream: one 4-bit group
    per read+write).  HBM GB/s per pass = 2 x state bytes x passes / time."""
    import torch

    from qml_essentials_b200 import config
    from qml_essentials_b200.model import Model

    prev = config.get_precision()
    config.set_precision("complex64")

    def one(label):
        model = Model(n_qubits=n_qubits, n_layers=8, circuit_type="Hardware_Efficient")
        rng = np.random.default_rng(1000)
        params = rng.uniform(0.0, 2 * np.pi, (1, *model._params_shape))
        inputs = np.array([[0.5]])
        spy = _Spy(ex)
        model.script.executor = spy
        ev = model(params=params, inputs=inputs)  # plan + first run
        model.script.executor = None
        plan, host_args, _ = spy.last
        call = ex.stage(plan, host_args, 1)
        call.launch()
        torch.cuda.synchronize()
        times = []
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = call.launch()
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
        ms = float(np.mean(times))
        handle = call.h
        state_bytes = 8 * 2 ** n_qubits
        passes = handle.n_passes
        res = out.cpu().numpy().reshape(-1)
        rec = {"strategy": handle.strategy, "passes": passes,
               "device_ops": handle.n_device_ops, "ms_per_circuit": ms,
               "ms_per_pass": ms / passes,
               "achieved_gbs": 2 * passes * state_bytes / (ms * 1e-3) / 1e9,
               "kernel": {4: "k_fstream<float> (2^13-amplitude tiles, TMA bulk copies, CX "
                             "folded into the frame)",
                          2: "k_stream<float,4,lean> (4-bit register groups, cp.async)"}.get(
                              handle.strategy, label),
               "checks": {"abs_expval_le_1": bool(np.all(np.abs(res) <= 1 + 1e-4)),
                          "repeatable": bool(np.allclose(res, np.asarray(ev).reshape(-1),
                                                         atol=1e-6))}}
        del call, out
        torch.cuda.empty_cache()
        return rec, res

    try:
        tile, res_tile = one("tile passes")
        group = None
        old = os.environ.get("QMLB_FSTREAM")
        os.environ["QMLB_FSTREAM"] = "0"
        try:
            group, res_group = one("register-group passes")
            group["max_abs_diff_vs_tile_passes"] = float(np.abs(res_group - res_tile).max())
        except Exception as exc:  # noqa: BLE001
            group = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        finally:
            if old is None:
                os.environ.pop("QMLB_FSTREAM", None)
            else:
                os.environ["QMLB_FSTREAM"] = old
        state_bytes = 8 * 2 ** n_qubits
        return {"workload": f"Model({n_qubits},8,'Hardware_Efficient') complex64 expval, "
                            "1 circuit, state in HBM", "n_qubits": n_qubits,
                "state_gib": state_bytes / 2 ** 30, "bytes_per_pass": 2 * state_bytes,
                "tape_gates": 44 * n_qubits, **tile, "register_group_stream": group,
                "note": "a tile pass holds ~5.5 register-group sub-passes per read+write of the "
                        "state and is bounded by shared-memory bandwidth + FP32 FMA, not by "
                        "HBM: fewer bytes per circuit (16 instead of 77 passes) at a lower "
                        "byte rate; the register-group stream shows the byte rate of a pass "
                        "that does one group's worth of work"}
    finally:
        config.set_precision(prev)


def sharded_leg(world, rank, local_bits=30):
    """Qubit-sharded statevector over all ranks: Model(local_bits + log2(N), 8,
    'Hardware_Efficient') complex64, top log2(N) state bits = rank index, global<->local
    swaps as NCCL all_to_all_single, <Z_q> all-reduced.  Wall time between barriers."""
    import torch
    import torch.distributed as dist

    from qml_essentials_b200 import config
    from qml_essentials_b200.model import Model
    from qml_essentials_b200.sharded import ShardedExecutor

    g = int(np.log2(world))
    if 2 ** g != world:
        return {"skipped": "needs a power-of-two rank count"}
    n = local_bits + g
    prev = config.get_precision()
    config.set_precision("complex64")
    try:
        model = Model(n_qubits=n, n_layers=8, circuit_type="Hardware_Efficient")
        params = np.random.default_rng(1000).uniform(0.0, 2 * np.pi, (1, *model._params_shape))
        inputs = np.array([[0.5]])
        se = ShardedExecutor()
        model.script.executor = se
        model(params=params, inputs=inputs)  # plan, epoch programs, NCCL warm-up
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ev = np.asarray(model(params=params, inputs=inputs)).reshape(-1)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        res = {"workload": f"Model({n},8,'Hardware_Efficient') complex64 expval, qubit-sharded "
                           f"over {world} GPUs ({local_bits} local bits)", "n_qubits": n,
               "seconds": dt, **se.stats, "abs_expval_le_1": bool(np.all(np.abs(ev) <= 1 + 1e-4))}
        # parity on hardware: the same kind of circuit at n = 28, sharded over all ranks vs
        # evolved on ONE GPU (rank 0, streaming frame engine)
        try:
            m28 = Model(n_qubits=28, n_layers=8, circuit_type="Hardware_Efficient")
            p28 = np.random.default_rng(7).uniform(0.0, 2 * np.pi, (1, *m28._params_shape))
            m28.script.executor = ShardedExecutor()
            sh28 = np.asarray(m28(params=p28, inputs=inputs)).reshape(-1)
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                m28.script.executor = None
                one = np.asarray(m28(params=p28, inputs=inputs)).reshape(-1)
                res["max_abs_err_vs_single_gpu"] = float(np.abs(sh28 - one).max())
                res["parity_check"] = "n = 28, same circuit family, all ranks vs rank 0 alone"
            dist.barrier()
        except Exception as exc:  # noqa: BLE001
            res["parity_check_error"] = f"{type(exc).__name__}: {exc}"[:200]
        return res
    finally:
        config.set_precision(prev)


class _Spy:
    """Captures the (plan, host arguments, batch) of a Model call so the same launch can be
    replayed on device-resident arguments."""

    def __init__(self, inner):
        self.inner, self.last = inner, None

    def __getattr__(self, k):
        return getattr(self.inner, k)

    def execute(self, plan, host_args, batch, chunk=None, to_host=True):
        self.last = (plan, host_args, batch)
        return self.inner.execute(plan, host_args, batch, chunk, to_host)


def staged_ms(ex, model, reps=5, **call_kw):
    """Device time (CUDA events) of one qmlb_run over HBM-resident arguments."""
    import torch
    import warnings

    spy = _Spy(ex)
    model.script.executor = spy
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model(**call_kw)
    plan, host_args, batch = spy.last
    call = ex.stage(plan, host_args, batch)
    call.launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        call.launch()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    model.script.executor = None
    return float(np.mean(ts)), batch, call.h


def config_legs(ex, peak_f64_tf):
    """The other small-n BASELINE configs, device-resident (SURVEY 8(d) cfg 1, 3, 4).
    `achieved` uses the ALGORITHMIC (dense, unfused tape) flop counts of SURVEY 8(d)."""
    from qml_essentials_b200 import config
    from qml_essentials_b200.model import Model

    prev = config.get_precision()
    config.set_precision("complex128")
    out = {}
    try:
        rng = np.random.default_rng(1000)
        cases = {
            "config_1": (lambda: Model(2, 1, "Circuit_19"), lambda m: dict(
                params=rng.uniform(0, 2 * np.pi, (1, *m._params_shape)),
                inputs=np.linspace(-np.pi, np.pi, 1024).reshape(-1, 1)), 1040,
                "Model(2,1,'Circuit_19') expval, 1024 inputs x 1 param set", "k_reg<double,2>"),
            "config_3": (lambda: Model(6, 3, "Circuit_15"), lambda m: dict(
                params=rng.uniform(0, 2 * np.pi, (20000, *m._params_shape)),
                execution_type="state"), 135168,
                "Model(6,3,'Circuit_15') statevectors of 20 000 parameter sets "
                "(expressibility / Meyer-Wallach input)", "k_frame<double,256> (64 states per CTA)"),
            "config_4": (lambda: Model(8, 4, "Strongly_Entangling"), lambda m: dict(
                params=rng.uniform(0, 2 * np.pi, (1, *m._params_shape)),
                inputs=np.linspace(-np.pi, np.pi, 4096).reshape(-1, 1),
                noise_params={"Depolarizing": 0.01, "AmplitudeDamping": 0.02},
                execution_type="density"), 2.69e9,
                "Model(8,4,'Strongly_Entangling') noisy density matrices, batch 4096 "
                "(4 GiB written once)",
                "k_frame_ptm<double,512> (real Pauli coefficients, cluster of 4 CTAs, DSMEM)"),
        }
        for name, (mk_model, mk_kw, flop, what, kernel) in cases.items():
            try:
                import warnings

                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    m = mk_model()
                ms, batch, h = staged_ms(ex, m, **mk_kw(m))
                tf = flop * batch / (ms * 1e-3) / 1e12
                out[name] = {"workload": what, "evals_per_s": batch / (ms * 1e-3),
                             "device_ms": ms, "batch": batch, "strategy": h.strategy,
                             "steps_or_passes": h.n_passes, "kernel": kernel,
                             "roofline": {"bound": "fp64_fma", "achieved": tf,
                                          "peak": peak_f64_tf, "unit": "TFLOP/s",
                                          "frac": tf / peak_f64_tf if peak_f64_tf else None,
                                          "note": "algorithmic flops of the dense unfused tape "
                                                  "(SURVEY 8(d)); the kernels fold CX and fuse "
                                                  "channels, so they execute fewer"}}
            except Exception as exc:  # noqa: BLE001
                out[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    finally:
        config.set_precision(prev)
    return out


def hbm_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args, json_out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, params, inputs = workload()
    n_sample = params.shape[0]  # the full workload of the GPU arm (same config)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, _, sample = oracle_cpu_evals_per_s(params, inputs, n_sample, cores, repeats=1)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    B = inputs.shape[0] * n_sample
    value = B * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "circuit evals/sec (batched)", "value": value,
        "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "evals_per_step_per_gpu": B,
                   "note": "reference-faithful CPU restatement (no XLA): the reference needs "
                           "JAX, absent from this image"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, ...) also
    write to file descriptor 1, so the real stdout is kept aside for the JSON line and fd 1
    is pointed at stderr for everything else."""
    sys.stdout.flush()
    keep = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return keep


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="complex128", choices=["complex128", "complex64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly")
    ap.add_argument("--no-config-legs", action="store_true",
                    help="skip the config 1 / 3 / 4 device-resident legs")
    ap.add_argument("--gate-pass-qubits", type=int, default=-1,
                    help="n for the HBM gate-pass leg (default: 32 on 1 GPU, 30 per rank "
                         "otherwise; 0 disables)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, json_out)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from qml_essentials_b200 import config, script
    from qml_essentials_b200.backend import QMLB_C128

    config.set_precision(args.precision)
    ex = script.get_executor()
    model, params, inputs = workload(seed=1000 + rank)  # every rank: its own samples
    B_I, B_P = inputs.shape[0], params.shape[0]
    B = B_I * B_P
    n_freq = B_I

    # ---- plan + stage (device-resident arguments) --------------------------------
    import ctypes as C

    spy = _Spy(ex)
    model.script.executor = spy
    model(params=params, inputs=inputs)  # records + compiles the plan, creates the program
    model.script.executor = None
    plan, host_args, _ = spy.last
    call = ex.stage(plan, host_args, B)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    # FCC statistics of the frequencies the circuit can carry (0 .. degree // 2): what a
    # batch-sharded Fourier fingerprint has to combine across ranks
    K = model.degree[0] // 2 + 1
    rows = torch.arange(K, dtype=torch.int32, device=dev)
    n_stat = 2 * (2 * K + K * K)  # complex128 as pairs of doubles
    coef = torch.empty((B_I, B_P), dtype=torch.complex128 if args.precision == "complex128"
                       else torch.complex64, device=dev)
    moments = torch.empty(n_stat, dtype=torch.float64, device=dev)
    reduced = torch.empty(n_stat, dtype=torch.float64, device=dev)
    dt_code = 1 if args.precision == "complex128" else 0
    lib = ex.lib
    collective = "none (1 GPU)"
    peer_ptrs = None
    if world > 1:
        try:
            import torch.distributed._symmetric_memory as symm

            nbytes = int(lib.qmlb_allreduce_buffer_bytes(n_stat))
            sbuf = symm.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
            sbuf.zero_()
            hdl = symm.rendezvous(sbuf, dist.group.WORLD)
            torch.cuda.synchronize()
            dist.barrier()
            peer_ptrs = (C.c_void_p * world)(*[int(q) for q in hdl.buffer_ptrs])
            collective = ("one-shot all-reduce over symmetric memory (k_allreduce_oneshot) on a "
                          "forked stream: step k publishes the statistics of step k - 1 while "
                          "k_reg runs; the last two reductions are drained inside the timed "
                          "region")
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] symmetric memory unavailable ({exc}); using NCCL", file=sys.stderr)
            collective = "NCCL all_reduce"

    side = torch.cuda.Stream(device=dev) if world > 1 else None

    def step_device():
        main = torch.cuda.current_stream(dev)
        st = main.cuda_stream
        rc = 0
        if world > 1 and peer_ptrs is not None:
            # two-deep pipeline: the statistics of step k - 1 are published (and those of step
            # k - 2 summed) on a side stream WHILE k_reg of step k runs - fork here, join before
            # k_coef_moments overwrites `moments`; the sequence is drained inside the timed
            # region (finish() below)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                rc |= lib.qmlb_allreduce_peer(peer_ptrs, world, rank, n_stat, moments.data_ptr(),
                                              reduced.data_ptr(), 1, side.cuda_stream)
        out = call.launch()  # (B, 4) expvals, flat order b = i * B_P + p
        rc |= lib.qmlb_grid_dft(out.data_ptr(), dt_code, B_I, B_P, N_QUBITS, None,
                                coef.data_ptr(), st)
        if world > 1 and peer_ptrs is not None:
            main.wait_stream(side)
        rc |= lib.qmlb_coef_moments(coef.data_ptr(), dt_code, rows.data_ptr(), K, B_P,
                                    moments.data_ptr(), st)
        if world > 1 and peer_ptrs is None:
            reduced.copy_(moments)
            dist.all_reduce(reduced)
        if rc != 0:
            raise RuntimeError(lib.qmlb_last_error().decode())
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    def drain():
        # publish the statistics of the last step, then sum them: `reduced` = their all-reduce
        if world > 1 and peer_ptrs is not None:
            st = torch.cuda.current_stream(dev).cuda_stream
            for mode in (1, 2):
                if lib.qmlb_allreduce_peer(peer_ptrs, world, rank, n_stat, moments.data_ptr(),
                                           reduced.data_ptr(), mode, st) != 0:
                    raise RuntimeError(lib.qmlb_last_error().decode())

    collective_err = None
    if world > 1 and peer_ptrs is not None:  # the one-shot all-reduce against NCCL's
        drain()
        want = moments.clone()
        dist.all_reduce(want)
        torch.cuda.synchronize()
        collective_err = float((reduced - want).abs().max())

    # The step is launch-bound at this size (k_pre x2 + k_reg + the statistics kernels + one
    # small NCCL all-reduce): capture it once in a CUDA graph and replay it.
    eager_step, graph, launches_per_step = step_device, None, None
    if not args.no_graph:
        try:
            l0 = ex.launch_count()
            warm = torch.cuda.Stream()
            warm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(warm):
                step_device()
            torch.cuda.current_stream().wait_stream(warm)
            graph = torch.cuda.CUDAGraph()
            l0 = ex.launch_count()
            with torch.cuda.graph(graph):
                graph_out = step_device()
            launches_per_step = ex.launch_count() - l0
            graph.replay()
            torch.cuda.synchronize()

            def step_device():  # noqa: F811
                graph.replay()
                return graph_out
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed ({exc}); running eagerly", file=sys.stderr)
            graph, step_device = None, eager_step
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ex.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for i, (s, e) in enumerate(ev):
        flush.fill_(1)  # L2 flush between timed iterations (untimed)
        s.record()
        out = step_device()
        if i == len(ev) - 1:
            drain()  # the reduction of the last step, inside the timed region
        e.record()
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    launches = ex.launch_count() - launches0
    if graph is not None:
        launches = launches_per_step * args.steps  # replayed: counted once at capture

    # kernel-only duration of the dominant kernel (one launch per step)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    for s, e in kev:
        flush.fill_(1)
        s.record()
        call.launch()
        e.record()
    torch.cuda.synchronize()
    kern_ms = sum(s.elapsed_time(e) for s, e in kev) / args.steps

    # ---- end to end through Coefficients.get_spectrum (host arrays in, host array out) ----
    from qml_essentials_b200.coefficients import Coefficients

    model.params = params
    for _ in range(args.warmup):
        spec, freqs = Coefficients.get_spectrum(model, mfs=MFS, shift=True, trim=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        spec, freqs = Coefficients.get_spectrum(model, mfs=MFS, shift=True, trim=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    res = model(params=params, inputs=inputs)  # raw expvals for the CPU cross-check below

    # ---- HBM-streaming regime: GB/s per fused gate pass (every rank: its own state) ----
    gp = None
    gp_n = args.gate_pass_qubits if args.gate_pass_qubits >= 0 else (32 if world == 1 else 30)
    if gp_n > 0:
        del flush
        torch.cuda.empty_cache()
        try:
            gp = gate_pass_leg(ex, gp_n, 2, dev)
        except Exception as exc:  # noqa: BLE001  (e.g. not enough free HBM on a shared box)
            gp = {"error": f"{type(exc).__name__}: {exc}"[:300], "n_qubits": gp_n}
    sh = None
    if world > 1 and gp_n > 0:
        torch.cuda.empty_cache()
        try:
            # 8 GPUs: BASELINE configs[4] - n = 35, 32 local bits (32 GiB shard + exchange
            # buffer per GPU); fewer ranks: 30 local bits
            sh = sharded_leg(world, rank, 32 if world == 8 and gp_n >= 30 else min(gp_n, 30))
        except Exception as exc:  # noqa: BLE001
            sh = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    clocks = sampler.stop() if rank == 0 else None
    gp_ms = gp.get("ms_per_circuit", 0.0) if gp else 0.0

    # measured while every rank is still alive (the ranks leave right after the last collective)
    peak_tf = ex.fma_peak_tflops(args.precision) if rank == 0 else None
    cfg_legs = None
    if rank == 0 and not args.no_config_legs:
        torch.cuda.empty_cache()
        cfg_legs = config_legs(ex, ex.fma_peak_tflops("complex128"))

    t = torch.tensor([dev_ms, e2e_s * 1e3, gp_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, gp_ms = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        total_evals = B * world * args.steps
        value = total_evals / (dev_ms * 1e-3)
        e2e_value = total_evals / (e2e_ms * 1e-3)
        prec64 = args.precision == "complex128"
        achieved_tf = ALGO_FLOP_PER_EVAL * B / (kern_ms * 1e-3) / 1e12
        line = {
            "metric": "circuit evals/sec (batched)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "complex128 (f64)" if prec64 else "complex64 (f32)",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "evals_per_step_per_gpu": B, "cache": "L2 flushed (512 MiB write) between "
                "timed iterations", "parallelism": f"batch-sharded x{world}",
                "launch": "CUDA graph replay" if graph is not None else "eager",
                "step": "k_pre x2, k_reg, k_grid_dft, k_coef_moments"
                        + (", k_allreduce_oneshot" if peer_ptrs is not None else ""),
                "collective": collective, "stat_bytes": n_stat * 8,
                "collective_max_abs_err_vs_nccl": collective_err,
            },
            "e2e": {"value": e2e_value, "unit": "evals/s",
                    "h2d_bytes_per_step": int(params.nbytes + inputs.nbytes),
                    "d2h_bytes_per_step": int(B_I * B_P * (16 if prec64 else 8)),
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "Coefficients.get_spectrum(model, mfs=8, shift=True, trim=True) "
                           "with model.params = NumPy (1024, 5, 12); returns the host "
                           "coefficient array " + str(tuple(spec.shape))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "fp64_fma" if prec64 else "fp32_fma",
                "kernel": f"k_reg<{'double' if prec64 else 'float'},4>",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf,
                "traffic": 1667584 if prec64 else None,
                "traffic_source": "ncu --set full capture of this kernel, not measured in "
                                  "this run (profiles/r2_final_kreg_digest.txt: dram read "
                                  "1.67 MB, write 0): parameters + hoisted tables in; the "
                                  "8.6 MB of results stay in L2",
                "kernel_ms": kern_ms,
                "note": "achieved = algorithmic (unfused, dense) 26624 flop/eval x 270336 "
                        "evals / CUDA-event launch time; peak = qmlb_fma_peak measured on "
                        "this GPU (no FMA figure in MEASURED_PEAKS.json).  The kernel fuses "
                        "and hoists, so it EXECUTES ~4x fewer flops than the algorithmic "
                        "count (frac can exceed 1); its FP64 pipe is 29 % busy (ncu)",
            },
        }
        if gp is not None and "error" not in gp:
            peak, src = hbm_peak_gbs()
            agg = world * 2 * gp["passes"] * 8 * 2 ** gp["n_qubits"] / (gp_ms * 1e-3) / 1e9
            gp.update({"achieved_gbs": agg, "per_gpu_gbs": agg / world, "peak_gbs": peak,
                       "peak_source": src, "frac": agg / world / peak, "n_gpus": world,
                       "ms_per_circuit": gp_ms, "ms_per_pass": gp_ms / gp["passes"]})
            rg = gp.get("register_group_stream")
            if isinstance(rg, dict) and "achieved_gbs" in rg:
                rg["frac"] = rg["achieved_gbs"] / peak
        if gp is not None:
            line["gate_pass"] = gp
        if sh is not None:
            line["qubit_sharded"] = sh
        if cfg_legs:
            line.update(cfg_legs)
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cpu_v, cpu_ev, sample = oracle_cpu_evals_per_s(params, inputs, B_P, cores)
            err = float(np.abs(res - cpu_ev).max())
            line["cpu_baseline"] = {"value": cpu_v, "unit": "evals/s", "cores": cores,
                                    "kind": "port", "sample": sample,
                                    "max_abs_err_vs_gpu": err}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        # orderly shutdown: drop the captured graph, then the process group.  The step graph
        # no longer holds an NCCL collective (round 1 left through os._exit because
        # destroying the communicator under a live graph could block); a watchdog still
        # bounds the teardown so a stuck collective cannot hold the node.
        graph = None
        torch.cuda.synchronize()
        done = threading.Event()

        def _teardown():
            try:
                dist.barrier()
                dist.destroy_process_group()
            finally:
                done.set()

        t_down = threading.Thread(target=_teardown, daemon=True)
        t_down.start()
        if not done.wait(60):
            print("[bench] process-group teardown timed out; leaving", file=sys.stderr)
            sys.stderr.flush()
            os._exit(0)


if __name__ == "__main__":
    main()
