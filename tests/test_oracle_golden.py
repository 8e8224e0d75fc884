"""CPU: pin oracle/restate.py against the fixtures produced by the UNMODIFIED reference (oracle/gen_golden.py).

Tolerance: the restatement and the reference run the same torch-CPU fp32 layer arithmetic but with different op
fusion / thread counts; 1e-5 normalised max error for tensors (tiny-magnitude biases see cancellation), 1e-6 for losses.  Integer state must match exactly."""
import numpy as np
import torch

from oracle import qp, restate
from tests.util import BS, NCLS, batches, load_golden, loaders, rel_err, tiny_model

TOL = 1e-5


def test_finetune_matches_reference():
    g = load_golden("finetune")
    for tag, f in g.items():
        m = tiny_model(f["init"])
        ld, sizes = loaders(f["data"])
        tr = restate.Trainer(m, "sgd", f["lr"], wd=f["wd"])
        best, log, _ = tr.train_model(ld, sizes, f["epochs"])
        assert abs(best - f["best_acc"]) < 1e-12
        ref_train = f["losses"]
        # the reference's recorded criterion calls interleave train and val batches; compare the train ones
        nb_t, nb_v = len(ld["train"]), len(ld["val"])
        ref = [l for e in range(f["epochs"]) for l in ref_train[e * (nb_t + nb_v): e * (nb_t + nb_v) + nb_t]]
        assert np.allclose(tr.batch_losses, ref, rtol=1e-6, atol=0)
        for k, v in f["final"].items():
            assert rel_err(m.state_dict()[k], v) <= TOL, k


def _run_penalty(which):
    g = load_golden(which)
    m = tiny_model(g["init"])
    names = [n for n, _ in m.named_parameters()]
    omega_acc = None
    for rnd in g["rounds"]:
        xp, yp = rnd["prev_data"]
        bl = batches(xp, yp)
        new = restate.fisher_pass(m, bl, len(xp)) if which == "ewc" else restate.mas_pass(m, bl)
        omega_acc = new if omega_acc is None else restate.accumulate_protocol(omega_acc, new)
        # from round 2 on the (round-1) head has no reg_params entry in the reference: it is never registered
        tracked = [n in rnd["reg_after_pass"] for n in names]
        for n, o, p, tr_ in zip(names, omega_acc, m.parameters(), tracked):
            if tr_:
                assert rel_err(o, rnd["reg_after_pass"][n]["omega"]) <= TOL, (which, n)
                assert rel_err(p.data, rnd["reg_after_pass"][n]["init_val"]) <= TOL
        reg = [dict(omega=o.clone(), init_val=p.data.clone()) if tr_ else None
               for o, p, tr_ in zip(omega_acc, m.parameters(), tracked)]
        m.classifier._modules["4"].load_state_dict(rnd["new_head"])
        reg[-1] = reg[-2] = None                 # fresh head is not in reg_params
        ld, sizes = loaders(rnd["data"])
        tr = restate.Trainer(m, "penalty", rnd["lr"], reg=reg, lam=rnd["lam"], wd=rnd["wd"])
        best, log, _ = tr.train_model(ld, sizes, rnd["epochs"])
        assert abs(best - rnd["best_acc"]) < 1e-12
        for k, v in rnd["final"].items():
            assert rel_err(m.state_dict()[k], v) <= TOL, (which, k)
        # stale head entries keep their omega in the reference dict; the accumulated omega continues with them
    return True


def test_ewc_matches_reference():
    assert _run_penalty("ewc")


def test_mas_matches_reference():
    assert _run_penalty("mas")


def test_si_matches_reference():
    g = load_golden("si")
    m = tiny_model(g["init"])
    names = [n for n, _ in m.named_parameters()]
    reg = None
    for r, rnd in enumerate(g["rounds"]):
        if r == 0:
            reg = [dict(omega=torch.zeros_like(p), w=torch.zeros_like(p), init_val=p.data.clone()) for p in m.parameters()]
        else:
            m.classifier._modules["4"].load_state_dict(rnd["head"])
            params = list(m.parameters())
            for i, p in enumerate(params):
                if i >= len(params) - 2:
                    reg[i] = dict(omega=torch.zeros_like(p), w=torch.zeros_like(p), init_val=p.data.clone())
                else:
                    o, w, ts = restate.si_consolidate(reg[i]["omega"], reg[i]["w"], p.data, reg[i]["init_val"])
                    reg[i] = dict(omega=o, w=w, init_val=ts)
        for n, rg in zip(names, reg):
            for key in ("omega", "w", "init_val"):
                assert rel_err(rg[key], rnd["reg_before"][n][key]) <= TOL, (r, n, key)
        ld, sizes = loaders(rnd["data"])
        tr = restate.Trainer(m, "si", rnd["lr"], reg=reg, lam=rnd["lam"])
        best, log, _ = tr.train_model(ld, sizes, rnd["epochs"])
        assert abs(best - rnd["best_acc"]) < 1e-12
        assert len(tr.batch_losses) == (rnd["epochs"] + 1) * len(ld["train"])     # the num_epochs + 1 quirk
        for k, v in rnd["final"].items():
            assert rel_err(m.state_dict()[k], v) <= TOL, k
        for n, rg in zip(names, reg):
            assert rel_err(rg["w"], rnd["reg_after"][n]["w"]) <= 1e-3, (r, n)   # w = -sum (theta_new - theta_old)*g0: the fp32 difference of nearly equal thetas carries ~1e-3 relative rounding noise per step in the REFERENCE arithmetic itself


def test_gem_matches_reference():
    g = load_golden("gem")
    base = tiny_model(g["init"], dropout=True)
    # wrap like gem.Net.__init__: 15-wide head, first 5 rows copied
    import torch.nn as nn
    head = nn.Linear(32, NCLS * g["n_tasks"])
    base.classifier._modules["6"] = head
    base.load_state_dict(g["wrapped_init"])
    store = {}
    fetch = lambda keys: torch.stack([store[k] for k in keys])
    gem = restate.GemOracle(base, g["n_tasks"], g["n_mem"], [NCLS] * g["n_tasks"], g["lr"], g["margin"], g["bs"], fetch)
    si = 0
    for t, (x, y) in enumerate(g["data"]):
        for b in range(3):
            st = g["steps"][si]
            si += 1
            xb, yb = x[b * 16:(b + 1) * 16], y[b * 16:(b + 1) * 16]
            for k, xi in zip(st["keys"], xb):
                store[k] = xi
            # dropout unit masks: the ones the reference drew (its DataLoader iterators also consume the host RNG,
            # so replaying the generator state alone does not reproduce them)
            loss, corr, stats = gem.observe(xb, t, yb, st["keys"], masks=st["masks"])
            assert stats["violations"] == st["violations"], (si, stats, st["violations"])
            assert gem.mem_cnt == st["mem_cnt"]
            assert abs(loss - st["loss"]) <= 1e-6 * max(1, abs(st["loss"]))
            assert corr == st["correct"]
            flat = torch.cat([p.data.reshape(-1) for p in base.parameters()])
            assert rel_err(flat, st["params"]) <= 5e-6, si
    assert torch.equal(gem.memory_labels, g["memory_labels"])
    for t in range(g["n_tasks"]):
        assert gem.exemplars[t] == g["exemplars"][t]
    assert rel_err(gem.grads, g["grads"]) <= 1e-5


def test_qp_known_answers():
    for c in load_golden("qp"):
        M, gvec = c["M"].numpy(), c["g"].numpy()
        x, v = qp.project2cone2(gvec.astype(np.float32), M.astype(np.float32), c["margin"])
        P = M.astype(np.float32).astype(np.float64)
        # v was generated from fp64 inputs; recompute on the same fp64 inputs for the known-answer check
        k = M.shape[0]
        Pm = M @ M.T
        Pm = 0.5 * (Pm + Pm.T) + 1e-3 * np.eye(k)
        v64 = qp.solve_lower_bounded_qp(Pm, -(M @ gvec), np.full(k, c["margin"]))
        assert np.abs(v64 - c["v"].numpy()).max() <= 1e-9 * max(1.0, np.abs(v64).max())
        assert (v64 >= c["margin"] - 1e-9).all()


def test_ragged_tails_match_reference():
    """Ragged last batches (56 = 16+16+16+8 importance images, 41 = 16+16+9 training, 18 = 16+2 validation): MAS's running
    average divides by the CURRENT batch size (train_MAS.py:168-173), mean-CE of a short batch, epoch statistics."""
    g = load_golden("ragged")
    f = g["finetune"]
    m = tiny_model(f["init"])
    ld, sizes = loaders(f["data"])
    tr = restate.Trainer(m, "sgd", f["lr"], wd=f["wd"])
    best, _, _ = tr.train_model(ld, sizes, f["epochs"])
    assert abs(best - f["best_acc"]) < 1e-12
    nb_t, nb_v = len(ld["train"]), len(ld["val"])
    assert (nb_t, nb_v) == (3, 2)
    ref = [l for e in range(f["epochs"]) for l in f["losses"][e * (nb_t + nb_v): e * (nb_t + nb_v) + nb_t]]
    assert np.allclose(tr.batch_losses, ref, rtol=1e-6, atol=0)
    for k, v in f["final"].items():
        assert rel_err(m.state_dict()[k], v) <= TOL, k
    for which in ("ewc", "mas"):
        r = g[which]
        m = tiny_model(r["init"])
        xp, yp = r["prev_data"]
        bl = batches(xp, yp)
        assert [len(b[0]) for b in bl] == [16, 16, 16, 8]
        om = restate.fisher_pass(m, bl, len(xp)) if which == "ewc" else restate.mas_pass(m, bl)
        names = [n for n, _ in m.named_parameters()]
        for n, o, p in zip(names, om, m.parameters()):
            assert rel_err(o, r["reg_after_pass"][n]["omega"]) <= TOL, (which, n)
            assert rel_err(p.data, r["reg_after_pass"][n]["init_val"]) <= TOL
        reg = [dict(omega=o.clone(), init_val=p.data.clone()) for o, p in zip(om, m.parameters())]
        m.classifier._modules["4"].load_state_dict(r["new_head"])
        reg[-1] = reg[-2] = None
        ld, sizes = loaders(r["data"])
        tr = restate.Trainer(m, "penalty", r["lr"], reg=reg, lam=r["lam"], wd=r["wd"])
        best, _, _ = tr.train_model(ld, sizes, r["epochs"])
        assert abs(best - r["best_acc"]) < 1e-12
        for k, v in r["final"].items():
            assert rel_err(m.state_dict()[k], v) <= TOL, (which, k)


def test_imm_precision_and_merges_match_reference():
    """Groundwork for SURVEY 8f-3: mode-IMM precision (sampled labels; divisor = #batches of the phase) and the mean /
    mode merges of three task models against methods/IMM/merge.py run unmodified."""
    g = load_golden("imm")
    precisions = []
    for t, state in enumerate(g["states"]):
        m = tiny_model(state)
        (xt, yt), (xv, yv) = g["data"][t]
        torch.manual_seed(g["seeds"][t])
        prec = restate.imm_precision_pass(m, {"train": batches(xt, yt), "val": batches(xv, yv)}, g["head_names"])
        assert set(prec) == set(g["precisions"][t])
        for n, v in g["precisions"][t].items():
            assert rel_err(prec[n], v) <= TOL, (t, n)
        precisions.append(prec)
    sums = [precisions[0]]
    for t in (1, 2):
        sums.append({n: sums[-1][n] + precisions[t][n] for n in precisions[t]})
    for i, upto in enumerate((1, 2)):
        mean = restate.imm_merge(g["states"], upto, g["head_names"])
        mode = restate.imm_merge(g["states"], upto, g["head_names"], precisions, sums[upto])
        intended = restate.imm_merge(g["states"], upto, g["head_names"], as_reference=False)
        for k in g["merged_mean"][i]:
            assert torch.equal(mean[k], g["merged_mean"][i][k]), ("mean", upto, k)      # the reference's no-op (see restate)
            assert rel_err(mode[k], g["merged_mode"][i][k]) <= TOL, ("mode", upto, k)
        k = "features.0.weight"
        assert rel_err(intended[k], sum(g["states"][j][k] for j in range(upto + 1)) / (upto + 1)) <= 1e-6
        assert not torch.equal(intended[k], mean[k])


def test_long_run_epoch_protocol_matches_reference():
    """lr cut at val_beat_counts == 5, stop at > 10 (SI: >= 10, epoch range num_epochs + 1): number of batches processed,
    final lr and best accuracy of 30-epoch runs of the reference's train_model with a vanishing learning rate."""
    g = load_golden("schedule")
    ld, sizes = loaders(g["data"])
    per_epoch = len(ld["train"]) + len(ld["val"])
    for which, kind in (("sgd", "sgd"), ("ewc", "penalty"), ("si", "si"), ("sgd_short", "sgd")):
        r = g[which]
        m = tiny_model(g["init"])
        reg = None
        if kind != "sgd":
            reg = [dict(omega=torch.ones_like(p) if kind == "penalty" else torch.zeros_like(p), init_val=p.data.clone(),
                        w=torch.zeros_like(p)) for p in m.parameters()]
        tr = restate.Trainer(m, kind, r["lr"], reg=reg, lam=1.0)
        best, log, _ = tr.train_model(ld, sizes, r["epochs"])
        n_calls = sum(len(ld[phase]) for _, phase, _, _ in log)
        assert n_calls == r["n_criterion_calls"], (which, n_calls, r["n_criterion_calls"])
        assert n_calls % per_epoch == 0
        assert abs(tr.lr - r["final_lr"]) <= 1e-12 * r["final_lr"], (which, tr.lr, r["final_lr"])
        assert abs(best - r["best_acc"]) < 1e-12
    assert g["sgd"]["n_criterion_calls"] == 12 * per_epoch and g["si"]["n_criterion_calls"] == 11 * per_epoch
    # divergence (lr = 1e3): EWC / SI abort after the first training phase (epoch loss > 1e4 or NaN), Finetune carries on
    for which, kind in (("diverge_ewc", "penalty"), ("diverge_si", "si"), ("diverge_sgd", "sgd")):
        r = g[which]
        m = tiny_model(g["init"])
        reg = None
        if kind != "sgd":
            reg = [dict(omega=torch.ones_like(p) if kind == "penalty" else torch.zeros_like(p), init_val=p.data.clone(),
                        w=torch.zeros_like(p)) for p in m.parameters()]
        tr = restate.Trainer(m, kind, r["lr"], reg=reg, lam=1.0)
        best, log, _ = tr.train_model(ld, sizes, r["epochs"])
        n_calls = sum(len(ld[phase]) for _, phase, _, _ in log)
        assert n_calls == r["n_criterion_calls"], (which, n_calls, r["n_criterion_calls"])
        assert abs(best - r["best_acc"]) < 1e-12, (which, best, r["best_acc"])
    assert g["diverge_ewc"]["n_criterion_calls"] == len(ld["train"]) and g["diverge_sgd"]["n_criterion_calls"] == 4 * per_epoch


def test_wide_fixture_matches_reference():
    """tests/golden/wide.pt (EWC + SI on a 64 / 64 / 128-channel VGGSlim, the layers the tensor-core kernels take): the
    restatement against the unmodified reference, and the fixture's own claim that no ReLU / pool decision of the run
    sits within 1e-5 of its boundary (found by a seed search, oracle/gen_golden.py MarginMonitor)."""
    from tests.util import WIDE_BS, wide_model
    g = load_golden("wide")
    assert g["ewc"]["min_margin"] >= 1.5e-5 and g["si"]["min_margin"] >= 1.5e-5
    r = g["ewc"]
    m = wide_model(r["init"])
    names = [n for n, _ in m.named_parameters()]
    xp, yp = r["prev_data"]
    om = restate.fisher_pass(m, batches(xp, yp, WIDE_BS), len(xp))
    for n, o in zip(names, om):
        assert rel_err(o, r["reg_after_pass"][n]["omega"]) <= TOL, n
    reg = [dict(omega=o.clone(), init_val=p.data.clone()) for o, p in zip(om, m.parameters())]
    m.classifier._modules["4"].load_state_dict(r["new_head"])
    reg[-1] = reg[-2] = None
    ld, sizes = loaders(r["data"], WIDE_BS)
    tr = restate.Trainer(m, "penalty", r["lr"], reg=reg, lam=r["lam"], wd=r["wd"])
    best, _, _ = tr.train_model(ld, sizes, r["epochs"])
    assert abs(best - r["best_acc"]) < 1e-12
    for k, v in r["final"].items():
        assert rel_err(m.state_dict()[k], v) <= TOL, k
    r = g["si"]
    m = wide_model(r["init"])
    reg = [dict(omega=r["reg_before"][n]["omega"].clone(), w=torch.zeros_like(p), init_val=r["reg_before"][n]["init_val"].clone())
           for n, p in m.named_parameters()]
    ld, sizes = loaders(r["data"], WIDE_BS)
    tr = restate.Trainer(m, "si", r["lr"], reg=reg, lam=r["lam"])
    best, _, _ = tr.train_model(ld, sizes, r["epochs"])
    assert abs(best - r["best_acc"]) < 1e-12
    for k, v in r["final"].items():
        assert rel_err(m.state_dict()[k], v) <= TOL, k
    for n, rg in zip(names, reg):
        assert rel_err(rg["w"], r["reg_after"][n]["w"]) <= 1e-3, n
