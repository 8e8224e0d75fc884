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
NB = 16                         # distinct synthetic batches the timed steps rotate over


def workload_config(model_name):
    """`config` of the JSON line -- byte-identical in the GPU arm and in `--impl reference` (the driver compares them)."""
    return {"workload": "%s bs=%d MAS penalised-SGD training step (train_MAS.train_model loop body), 64x64 synthetic inputs, "
                        "BASELINE configs[2]" % (model_name, GLOBAL_BATCH),
            "global_batch": GLOBAL_BATCH,
            "l2": "per-step working set (activations + gradients, ~0.5 GB at batch 200) exceeds the 126 MB L2; inputs rotate "
                  "over %d distinct batches" % NB}


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
    """The reference's CPU path for this step: oracle/restate.py (the port; /root/reference is absent on the box) on the
    host cores.  torch's CPU convolutions do not scale to every thread count, so one step is timed at a few thread counts
    (all cores, then halvings) and the fastest is used.  Returns (images/sec, threads used, sample text, ms/step)."""
    from oracle import restate
    model = build_model(model_name)
    g = torch.Generator().manual_seed(11)
    reg = [dict(omega=torch.rand(p.shape, generator=g) * 1e-3, init_val=p.data.clone()) for p in model.parameters()]
    reg[-1] = reg[-2] = None
    tr = restate.Trainer(model, "penalty", 0.01, reg=reg, lam=3.0)
    model.train()
    xs, ys = synth_batches(min(NB, 4), batch, 5)
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
        tr.step(xs[i % len(xs)], ys[i % len(xs)])
    t0 = time.perf_counter()
    for i in range(steps):
        tr.step(xs[i % len(xs)], ys[i % len(xs)])
    dt = time.perf_counter() - t0
    sample = "%d steps of batch %d after %d warm-up (%s, penalised SGD, %d of %d host threads: fastest of a thread-count sweep)" % (
        steps, batch, warmup, model_name, best_n, ncpu)
    return batch * steps / dt, best_n, sample, dt / steps * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one CPU step is ~0.5 s: the full --steps / --warmup run stays within minutes up to a few hundred steps
    steps, warm = max(1, min(args.steps, 400)), max(0, min(args.warmup, 50))
    v, cores, sample, ms = cpu_reference_steps(args.model, steps, warm)
    line = {"impl": "reference", "metric": "images/sec/task", "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.model),
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference CPU path = oracle/restate.py (pinned to the unmodified reference by tests/test_oracle_golden.py) "
                    "on torch CPU; /root/reference does not exist on the GPU box"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- GPU arm
def measured_traffic():
    """DRAM bytes per step of the conv launches from the committed `ncu --set full` capture (profiles/r2_conv_traffic.json,
    written by tools/ncu_traffic.py from the same command)."""
    p = os.path.join(ROOT, "profiles", "r2_conv_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


def conv_roofline(eng, model, step_fn, steps, per, pk, mode_name, ms_graph=None):
    """Time every conv-stack call of `steps` eager steps with CUDA events on the launching stream (Engine._timed)."""
    conv_f, _ = conv_flops_per_image(model)
    eng.conv_events = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step_fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms_eager = e0.elapsed_time(e1) / steps
    conv_ms = sum(a.elapsed_time(b) for a, b in eng.conv_events) / steps
    n_calls = len(eng.conv_events) // steps
    eng.conv_events = None
    tf = conv_f * per / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    return {"bound": "tensor", "achieved": tf, "peak": pk["tensor_burst"], "unit": "TFLOP/s", "frac": tf / pk["tensor_burst"],
            "traffic": None,
            "kernel": "conv stack: fused first layer + pl::conv_planes_kernel fwd / dgrad / wgrad (+ split-K reduce, bias grad) -- "
                      "%d calls per step, %s" % (n_calls, mode_name),
            "algorithmic_gflop_per_step": conv_f * per / 1e9, "ms_per_step_in_kernel": conv_ms, "mma_passes": 3,
            "frac_of_sustained_peak": tf / pk["tensor"], "frac_of_split_ceiling": 3 * tf / pk["tensor_burst"],
            "share_of_step": conv_ms / ms_eager,
            "measured": "CUDA events around every conv-stack call over %d eager steps (%.3f ms/step eager%s)" % (
                steps, ms_eager, "" if ms_graph is None else ", %.3f ms/step as replayed graph" % ms_graph),
            "peak_source": pk["src"] + ", bf16 dense burst (the timed region is ~60 ms at full clocks)"}, ms_eager


def run_gpu_arm(args):
    # native libraries (NCCL prints its version banner) write to fd 1: keep the real stdout for the ONE JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    from clsurvey_b200 import _capi, dist as cdist
    from clsurvey_b200.data import PinnedLoader
    from clsurvey_b200.engine import Engine
    from clsurvey_b200.methods.MAS import train_MAS
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
    xs, ys = synth_batches(NB, GLOBAL_BATCH, 5)
    xs_d = [x[lo:hi].contiguous().to(dev) for x in xs]
    ys_d = [y[lo:hi].contiguous().to(dev) for y in ys]
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
    value = GLOBAL_BATCH * args.steps / (ms / 1e3)
    # N > 1: the same per-rank step without its collectives (every rank steps on its local gradient; train_model re-syncs the
    # replicas before the end-to-end region) -> compute time and exposed all-reduce time of the data-parallel step
    comp_ms = None
    if world > 1 and state["run"] is not None:
        import torch.distributed as td
        ok = 1
        try:
            with cdist.local_only():
                run_local = eng.graphed(("bench_local", per), per, body)
                for i in range(2):
                    run_local(xs_d[i % NB], ys_d[i % NB])
            torch.cuda.synchronize()
        except Exception as e:
            sys.stderr.write("local-only step graph failed: %r\n" % (e,))
            ok = 0
        flag = torch.tensor([ok], device=dev)
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        if int(flag.item()):
            comp_ms = timed(lambda i: run_local(xs_d[i % NB], ys_d[i % NB]), args.steps) / args.steps
    mode_name = {0: "fp32 FFMA (SIMT)", 1: "tf32x3 tcgen05", 2: "tf32x1 tcgen05",
                 3: "bf16x3 tcgen05: TMA-fed NHWC hi/lo planes (conv + hidden Linear layers) + tf32x3 (task head)"}[args.mm_mode]
    # per-kernel timing of the dominant (conv) launches: the same K steps run eagerly with CUDA events around every conv
    # call (events cannot be read back from inside a replayed graph)
    l0 = _capi.lib().clb_launch_count()
    roof, ms_eager = conv_roofline(eng, model, step_eager, args.steps, per, pk, mode_name, ms / args.steps)
    launches = (_capi.lib().clb_launch_count() - l0) // max(args.steps, 1) * args.steps
    tr = measured_traffic()
    if tr is not None and world == 1 and GLOBAL_BATCH == 200:
        roof["traffic"] = tr["bytes_per_step"]
        roof["traffic_note"] = tr["note"]
        roof["algorithmic_bytes_per_step"] = tr.get("algorithmic_bytes_per_step")
    # the sampler ran across the two back-to-back device-timed regions (graph replay, per-kernel eager pass): all under
    # load; it is stopped before the end-to-end region so that its nvidia-smi subprocesses never compete with host code
    clocks = sampler.stop() if sampler else None

    # ---- end to end: the reference-facing plugin call train_MAS.train_model (src/methods/MAS/train_MAS.py:208-335) on an
    # in-memory task in PINNED HOST memory: every step copies its mini-batch host -> device, the per-batch loss / #correct
    # are read back once per phase like the trainer does.  One epoch = n_e2e training batches + one validation batch; E2E_EPOCHS
    # epochs, so that the checkpoint files of the call (epoch.pth.tar at epoch 0, best_model.pth.tar per validation
    # improvement, train_MAS.py:207-227; 126 MB each) weigh as they do in a task-length run.
    n_e2e, E2E_EPOCHS = max(args.steps, 40), 10
    g2 = torch.Generator().manual_seed(6)
    xe = torch.randn(n_e2e * GLOBAL_BATCH, *IN_SHAPE, generator=g2)
    ye = torch.randint(0, NUM_CLASSES, (n_e2e * GLOBAL_BATCH,), generator=g2)
    train_ds = torch.utils.data.TensorDataset(xe, ye)
    val_ds = torch.utils.data.TensorDataset(xe[:GLOBAL_BATCH].clone(), ye[:GLOBAL_BATCH].clone())
    warm_ds = torch.utils.data.TensorDataset(xe[:4 * GLOBAL_BATCH].clone(), ye[:4 * GLOBAL_BATCH].clone())
    import tempfile
    tmp = tempfile.mkdtemp(prefix="clb_bench_")
    crit = torch.nn.CrossEntropyLoss()

    def plugin_call(ds, epochs=1):
        loaders = {"train": PinnedLoader(ds, GLOBAL_BATCH), "val": PinnedLoader(val_ds, GLOBAL_BATCH)}
        sizes = {"train": len(ds), "val": len(val_ds)}
        o = train_MAS.Weight_Regularized_SGD(model.parameters(), 0.01, momentum=0.9, weight_decay=0.0)
        with open(os.devnull, "w") as dn:
            old = sys.stdout
            sys.stdout = dn
            try:
                train_MAS.train_model(model, crit, o, 0.01, loaders, sizes, True, epochs, exp_dir=tmp + "/", resume="")
            finally:
                sys.stdout = old

    plugin_call(warm_ds)                                 # warm-up: graph capture for this configuration, pinned allocations
    plugin_call(warm_ds)
    PinnedLoader(train_ds, GLOBAL_BATCH)                 # the task's tensors are page-locked ONCE (data.py caches them per dataset
    #                                                      object): the timed call starts with its inputs in pinned host memory
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plugin_call(train_ds, E2E_EPOCHS)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as td
        t = torch.tensor([ms_e2e], device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_e2e = t.item()
    n_e2e_steps = n_e2e * E2E_EPOCHS
    e2e = GLOBAL_BATCH * n_e2e_steps / (ms_e2e / 1e3)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)

    # ---- north_star's conv-roofline case: VGG-11 on a [64, 3, 64, 64] input (233.7 GF fwd + bwd)
    roof64 = None
    if rank == 0 and world == 1 and not has_dropout:
        m64 = build_model(args.model)
        e64 = Engine(m64, IN_SHAPE, 64)
        m64.reg_params = {p: {"omega": torch.zeros_like(p.data), "init_val": p.data.clone()} for p in list(m64.parameters())[:-2]}
        m64.reg_params["lambda"] = 3.0
        o64 = Weight_Regularized_SGD(m64.parameters(), 0.01, momentum=0.9, weight_decay=0.0)
        m64.train()
        x64 = [x[:64].contiguous().to(dev) for x in xs]
        y64 = [y[:64].contiguous().to(dev) for y in ys]

        def step64(i):
            e64.fwd_loss_bwd(x64[i % NB], y64[i % NB], denom=64, train=True)
            o64.step(m64.reg_params)
        for i in range(3):
            step64(i)
        roof64, _ = conv_roofline(e64, m64, step64, args.steps, 64, pk, mode_name)
        del e64, m64, o64

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
                  "traffic": 12.0 * P, "traffic_note": "algorithmic = measured: ncu dram bytes 629-630 MB per launch vs 685 MB "
                  "algorithmic (the tail of omega is still dirty in L2 when the launch retires), profiles/r1_ncu_full_fisher.csv",
                  "kernel": "fisher_kernel (omega += g*g/N)", "bytes_per_param": 12, "params": P,
                  "ms_per_launch": fms, "peak_source": pk["src"]}
        del om, gr

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_steps(args.model, 5, 1)
        cpu = {"value": v, "unit": "images/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "images/sec/task", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {0: "f32", 1: "tf32x3->f32", 2: "tf32", 3: "bf16x3/tf32x3->f32"}[args.mm_mode],
            "data": "synthetic", "config": workload_config(args.model),
            "details": {"per_gpu_batch": per, "parallelism": "dp%d" % world, "matmul_mode": mode_name,
                        "launch": "cuda graph replay (1 graph launch = %d kernels)" % (launches // max(args.steps, 1)) if state["run"] else "eager"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(per * (IN_SHAPE[0] * IN_SHAPE[1] * IN_SHAPE[2] * 4 + 8)),
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / n_e2e_steps, "steps": n_e2e_steps,
                    "api": "train_MAS.train_model(model, criterion, Weight_Regularized_SGD, lr, loaders, sizes, True, %d epochs, ...) on "
                           "an in-memory task already page-locked in host memory: %d training batches + 1 validation batch per epoch, per-phase loss read-back, "
                           "epoch.pth.tar + best_model.pth.tar checkpoints written (background writer)" % (E2E_EPOCHS, n_e2e)},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_n64": roof64, "roofline_fisher": fisher, "cpu_baseline": cpu,
        }
        if comp_ms is not None:                              # data-parallel step = per-rank compute + exposed all-reduce
            line["details"]["compute_ms_per_step"] = comp_ms
            line["details"]["exposed_allreduce_ms_per_step"] = max(ms / args.steps - comp_ms, 0.0)
            line["details"]["conv_ms_per_step_timed_call_by_call"] = roof.get("ms_per_step_in_kernel") if roof else None
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
