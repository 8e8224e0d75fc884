"""bench.py -- images/sec/task of the hot path (VGG-11, TinyImagenet-shaped 64x64 inputs, batch 200, MAS-style penalised
SGD step) on N B200s, plus the Fisher/Omega accumulator bandwidth, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--model VGG11_cl_512_512]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one training mini-batch of the per-task loop (src/methods/MAS/train_MAS.py:208-335): forward, mean-CE,
backward, [gradient all-reduce,] fused penalised SGD update -- all through the C ABI (include/clb.h).
`value` times that with inputs resident in HBM; `e2e` times the same step driven from pinned HOST buffers with the
H2D copy of the batch and the D2H read of the loss inside the timed region.  Synthetic data (N(0,1) images, uniform
labels), random-init weights, random positive Omega: there is no dataset / checkpoint on the box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

GLOBAL_BATCH = 200
IN_SHAPE = (3, 64, 64)
NUM_CLASSES = 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, src="fallback (B200_PROFILING.md)")


def conv_flops_per_image(model, train=True):
    """Algorithmic FLOPs of the conv/linear stack: 2*N*K*C*R*S*P*Q per pass; fwd + wgrad + dgrad (no dgrad for the
    first layer) -- SURVEY.md 8d."""
    import torch.nn as nn
    H, W = IN_SHAPE[1:]
    conv, lin, first = 0.0, 0.0, True
    for m in model.features:
        if isinstance(m, nn.Conv2d):
            P = (H + 2 * m.padding[0] - m.kernel_size[0]) // m.stride[0] + 1
            Q = (W + 2 * m.padding[1] - m.kernel_size[1]) // m.stride[1] + 1
            f = 2.0 * m.out_channels * m.in_channels * m.kernel_size[0] * m.kernel_size[1] * P * Q
            conv += f * ((2 if first else 3) if train else 1)
            first = False
            H, W = P, Q
        elif isinstance(m, nn.MaxPool2d):
            k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0]
            s = m.stride if isinstance(m.stride, int) else m.stride[0]
            H, W = (H - k) // s + 1, (W - k) // s + 1
    for m in model.classifier:
        if isinstance(m, nn.Linear):
            lin += 2.0 * m.in_features * m.out_features * (3 if train else 1)
    return conv, lin


class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 8:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def build_model(name):
    from clsurvey_b200.models import parse_model_name
    torch.manual_seed(7)                       # utils.set_random(7) before model creation
    return parse_model_name(name, IN_SHAPE[1:], NUM_CLASSES)


def synth_batches(n_batches, per, seed):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(per, *IN_SHAPE, generator=g) for _ in range(n_batches)]
    ys = [torch.randint(0, NUM_CLASSES, (per,), generator=g) for _ in range(n_batches)]
    return xs, ys


# ------------------------------------------------------------------------------------------------- CPU arms
def cpu_reference_steps(model_name, steps, warmup, batch=GLOBAL_BATCH):
    """The reference's CPU path for this step: oracle/restate.py (the port; /root/reference is absent on the box),
    all host threads.  Returns (images/sec, cores, sample description)."""
    from oracle import restate
    model = build_model(model_name)
    g = torch.Generator().manual_seed(11)
    reg = [dict(omega=torch.rand(p.shape, generator=g) * 1e-3, init_val=p.data.clone()) for p in model.parameters()]
    reg[-1] = reg[-2] = None
    tr = restate.Trainer(model, "penalty", 0.01, reg=reg, lam=3.0)
    model.train()
    xs, ys = synth_batches(2, batch, 5)
    # "all the host threads it can use": torch's CPU conv does not scale to every thread count, so time one step at a
    # few thread counts (all cores, then halvings) and keep the fastest for the measured sample
    ncpu = os.cpu_count() or 1
    best_t, best_n = None, ncpu
    for n in sorted({ncpu, max(ncpu // 2, 1), max(ncpu // 4, 1), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
        torch.set_num_threads(n)
        tr.step(xs[0], ys[0])
        t0 = time.perf_counter()
        tr.step(xs[1], ys[1])
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, best_n = dt, n
    torch.set_num_threads(best_n)
    for i in range(warmup):
        tr.step(xs[i % 2], ys[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        tr.step(xs[i % 2], ys[i % 2])
    dt = time.perf_counter() - t0
    return batch * steps / dt, torch.get_num_threads(), "%d steps of batch %d (%s, penalised SGD) after %d warm-up" % (
        steps, batch, model_name, warmup), dt / steps * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 8))          # bounded sample: one step is ~1-2 s of CPU work
    warm = max(1, min(args.warmup, 2))
    v, cores, sample, ms = cpu_reference_steps(args.model, steps, warm)
    line = {"impl": "reference", "metric": "images/sec/task", "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s bs=%d MAS-style penalised SGD step, 64x64 inputs" % (args.model, GLOBAL_BATCH),
                       "global_batch": GLOBAL_BATCH, "note": "reference CPU path = oracle/restate.py port on torch CPU"},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    # native libraries (NCCL prints its version banner) write to fd 1: keep the real stdout for the ONE JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    from clsurvey_b200 import _capi, dist as cdist
    from clsurvey_b200.engine import Engine
    from clsurvey_b200.methods.optim import Weight_Regularized_SGD
    _capi.lib()
    cdist.init()
    world, rank = cdist.world_size(), cdist.rank()
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    _capi.call("clb_set_matmul_mode", args.mm_mode)
    model = build_model(args.model)
    lo, hi = cdist.shard_rows(GLOBAL_BATCH)
    per = hi - lo
    eng = Engine(model, IN_SHAPE, max(per, 1))
    # MAS state: random positive omega, theta* = theta, fresh head unpenalised (main_MAS.py:72-80)
    g = torch.Generator().manual_seed(11)
    params = list(model.parameters())
    model.reg_params = {p: {"omega": (torch.rand(p.shape, generator=g) * 1e-3).to(dev), "init_val": p.data.clone()}
                        for p in params[:-2]}
    model.reg_params["lambda"] = 3.0
    opt = Weight_Regularized_SGD(model.parameters(), 0.01, momentum=0.9, weight_decay=0.0)
    model.train()
    NB = 16
    xs, ys = synth_batches(NB, GLOBAL_BATCH, 5)
    xs_h = [x[lo:hi].contiguous().pin_memory() for x in xs]
    ys_h = [y[lo:hi].contiguous().pin_memory() for y in ys]
    xs_d = [x.to(dev) for x in xs_h]
    ys_d = [y.to(dev) for y in ys_h]
    conv_f, lin_f = conv_flops_per_image(model)
    pk = peaks()

    def body(xb, yb):
        eng.fwd_loss_bwd(xb, yb, denom=GLOBAL_BATCH, train=True, dp_overlap=not args.no_overlap)
        opt.step(model.reg_params)

    has_dropout = any(op["kind"] == "dropout" for op in eng.ops)
    eng.dropout_rng = "device"                          # AlexNet: masks drawn on the device (parity tests use host masks)
    use_graph = (not args.no_graph) and (world == 1 or not args.no_graph_dp) and not has_dropout
    state = {"run": None}

    def step_eager(i):
        body(xs_d[i % NB], ys_d[i % NB])

    def step_dev(i):                                     # `value`: inputs already resident in HBM
        if state["run"] is None:
            return step_eager(i)
        state["run"](xs_d[i % NB], ys_d[i % NB])         # device->device copy into the graph's static buffers + replay

    xbuf = torch.empty_like(xs_d[0])
    ybuf = torch.empty_like(ys_d[0])

    def step_e2e(i):                                     # `e2e`: host buffers, H2D + D2H inside the timed region
        if state["run"] is None:
            xbuf.copy_(xs_h[i % NB], non_blocking=True)
            ybuf.copy_(ys_h[i % NB], non_blocking=True)
            body(xbuf, ybuf)
        else:
            state["run"](xs_h[i % NB], ys_h[i % NB])     # pinned host -> static device buffers (H2D) + graph replay
        return eng.loss_dev.item()                       # D2H read of the step's loss (a host sync, like the reference)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as td
            td.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as td
            t = torch.tensor([ms], device=dev)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            ms = t.item()
        return ms

    for i in range(args.warmup):
        step_eager(i)
    if use_graph:
        ok = 1
        try:
            state["run"] = eng.graphed(("bench", per), per, body)     # whole step = one CUDA-graph launch
        except Exception as e:                                        # e.g. an NCCL build that cannot be captured
            sys.stderr.write("graph capture failed, running eagerly: %r\n" % (e,))
            ok = 0
        if world > 1:                                                 # all ranks replay, or none does
            import torch.distributed as td
            flag = torch.tensor([ok], device=dev)
            td.all_reduce(flag, op=td.ReduceOp.MIN)
            ok = int(flag.item())
        if not ok:
            state["run"] = None
            eng.drop_graphs()
        for i in range(2):
            step_dev(i)
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(step_dev, args.steps)
    # per-kernel timing of the dominant (conv implicit-GEMM) launches: the same K steps run eagerly with CUDA events
    # around every conv launch (events cannot be read back from inside a replayed graph)
    l0 = _capi.lib().clb_launch_count()
    eng.conv_events = []
    ms_eager = timed(step_eager, args.steps)
    launches = (_capi.lib().clb_launch_count() - l0) // max(args.steps, 1) * args.steps
    conv_ms = sum(a.elapsed_time(b) for a, b in eng.conv_events) / max(args.steps, 1)
    n_conv_launch = len(eng.conv_events) // max(args.steps, 1)
    eng.conv_events = None
    value = GLOBAL_BATCH * args.steps / (ms / 1e3)
    # the sampler ran across the two back-to-back device-timed regions (graph replay, per-kernel eager pass): all under
    # load; it is stopped before the end-to-end region so that its nvidia-smi subprocesses never compete with host code
    clocks = sampler.stop() if sampler else None
    for i in range(min(args.warmup, 3)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    e2e = GLOBAL_BATCH * args.steps / (ms_e2e / 1e3)

    # Fisher / Omega accumulator bandwidth (AlexNet-sized flat buffer, config C2: P = 57,085,780 > L2)
    fisher = None
    if rank == 0:
        P = 57085780 // 4 * 4
        om = torch.zeros(P, device=dev)
        gr = torch.randn(P, device=dev)
        for _ in range(3):
            _capi.call("clb_fisher_accum", om.data_ptr(), gr.data_ptr(), 8000.0, P, torch.cuda.current_stream().cuda_stream)
        reps = 20
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            _capi.call("clb_fisher_accum", om.data_ptr(), gr.data_ptr(), 8000.0, P, torch.cuda.current_stream().cuda_stream)
        ev[1].record()
        torch.cuda.synchronize()
        fms = ev[0].elapsed_time(ev[1]) / reps
        gbs = 12.0 * P / (fms * 1e-3) / 1e9
        fisher = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                  "traffic": None, "kernel": "fisher_kernel (omega += g*g/N)", "bytes_per_param": 12, "params": P,
                  "ms_per_launch": fms, "peak_source": pk["src"]}
        del om, gr

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_steps(args.model, 5, 1)
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        conv_tf = conv_f * per / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        mode_name = {0: "fp32 FFMA (SIMT)", 1: "tf32x3 tcgen05", 2: "tf32x1 tcgen05",
                     3: "bf16x3 tcgen05 (conv fwd/dgrad/wgrad) + tf32x3 (rest)"}[args.mm_mode]
        line = {
            "metric": "images/sec/task", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {0: "f32", 1: "tf32x3->f32", 2: "tf32", 3: "bf16x3/tf32x3->f32"}[args.mm_mode],
            "data": "synthetic",
            "config": {"workload": "%s bs=%d MAS-style penalised SGD step, 64x64 inputs (BASELINE configs[2])" % (
                args.model, GLOBAL_BATCH), "global_batch": GLOBAL_BATCH, "per_gpu_batch": per,
                "parallelism": "dp%d" % world, "matmul_mode": mode_name,
                "launch": "cuda graph replay (1 graph launch = %d kernels)" % (launches // max(args.steps, 1)) if state["run"] else "eager",
                "l2": "per-step working set (activations+grads ~0.6 GB at batch 200) exceeds the 126 MB L2; inputs "
                      "rotate over %d distinct batches" % NB},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(xbuf.numel() * 4 + ybuf.numel() * 8),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": conv_tf, "peak": pk["tensor"], "unit": "TFLOP/s",
                         "frac": conv_tf / pk["tensor"], "traffic": None,
                         "kernel": "conv2d implicit-GEMM fwd+wgrad+dgrad (%d launches/step, %s)" % (n_conv_launch, mode_name),
                         "algorithmic_gflop_per_step": conv_f * per / 1e9, "ms_per_step_in_kernel": conv_ms,
                         # the fp32-parity split issues 3 MMAs per algorithmic one (bf16 rate in mode 3, tf32 = half of
                         # it in mode 1): fraction of the rate the tensor pipe can give to THIS arithmetic
                         "mma_passes": {0: 0, 1: 3, 2: 1, 3: 3}[args.mm_mode],
                         "frac_of_split_ceiling": (conv_tf * {0: 0, 1: 6, 2: 2, 3: 3}[args.mm_mode] / pk["tensor"]),
                         "share_of_step": conv_ms / (ms_eager / args.steps),
                         "measured": "CUDA events around every conv launch over %d eager steps (%.3f ms/step eager, "
                                     "%.3f ms/step as replayed graph)" % (args.steps, ms_eager / args.steps, ms / args.steps),
                         "peak_source": pk["src"] + ", bf16 dense sustained",
                         "traffic_note": "aggregate of all conv launches, so no single per-launch figure; ncu --set full "
                                         "per launch: 16-106 MB DRAM read + 0-59 MB written = the operand / output bytes "
                                         "(operands are L2-resident, no re-reads): profiles/r1_ncu_full_bf16*.csv"},
            "roofline_fisher": fisher,
            "cpu_baseline": cpu,
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    # NCCL keeps a communicator alive while a captured graph still references it: drop the graphs first, and never let
    # a slow teardown hold the job after the result line is out
    state["run"] = None
    eng.drop_graphs()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        import threading
        t = threading.Timer(20.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
    cdist.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="clb", choices=["clb", "reference"])
    ap.add_argument("--model", default="VGG11_cl_512_512")
    ap.add_argument("--mm-mode", type=int, default=int(os.environ.get("CLB_MM_MODE", "3")),
                    help="0 exact-fp32 FFMA, 1 tcgen05 TF32x3 (fp32 parity), 2 tcgen05 TF32x1 (fast, non-parity), "
                         "3 tcgen05 bf16x3 conv kernels + TF32x3 elsewhere (fp32 parity, default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--global-batch", type=int, default=200, help="diagnostics only: the BASELINE workload is 200")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-graph-dp", action="store_true", help="N > 1: launch eagerly instead of replaying the step (incl. the NCCL all-reduces) as one CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one all-reduce after backward instead of the overlapped tail/head pair")
    args = ap.parse_args()
    global GLOBAL_BATCH
    GLOBAL_BATCH = args.global_batch
    args.warmup = max(args.warmup, 3) if args.impl == "clb" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
