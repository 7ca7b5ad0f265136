#!/usr/bin/env python
"""bench.py — NSF-SVGP ELBO forward+backward steps/sec on B200 (BASELINE.json metric), 1..8 GPUs.

Default workload (BASELINE.json configs[1]): NSF2(SVGP(NSF_RBF)), N=32768 spots, M=1024 inducing points, L=10 factors,
G=2000 genes, E=1, fp32, synthetic Slide-seq-shaped data (gpzoo_b200.synthetic.nsf_problem).
One step = ELBO forward + backward producing every parameter gradient (+ the all-reduce of the shared-parameter gradients when
N>1); the optimiser update is excluded.

Multi-GPU: data parallel over spots, shared gradients summed with one NCCL all-reduce.  The headline `value` is STRONG scaling —
the 32768 spots of configs[1] split over the N ranks, i.e. the same job at every N, which is what BASELINE.json's north_star
quotes ("≥6x at 8 GPUs") and what the reference arm times — and the `weak` block of the same JSON line is the fixed-per-GPU-size
measurement (every GPU runs 32768 spots, global minibatch N x 32768).

`--config 3|4|5` times the other BASELINE.json configs (VNNGP K=8; MGGP-SVGP M=2048; hybrid M=4096; the last two with a fixed
minibatch per GPU, as they are defined) with the same step / timing code.  The driver's contract (`--gpus N --steps K --warmup W`,
no --config) always measures configs[1].

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU PyTorch path (oracle port) on the host cores
instead, on the same config, and adds the stock PyTorch-eager time of the same op sequence on this GPU (`eager_gpu`).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NSF-SVGP ELBO fwd+bwd steps/sec"
CONFIGS = {
    2: dict(kind="svgp", N=32768, M=1024, L=10, G=2000, E=1, coord_scale=100.0, lengthscale=1.7, jitter=0.1, seed=1,
            workload="NSF2(SVGP(NSF_RBF)) N=32768 M=1024 L=10 G=2000 E=1 (BASELINE.json configs[1])"),
    3: dict(kind="vnngp", N=4000, M=1000, L=10, G=2000, E=10, K=8, coord_scale=2.0, lengthscale=1.0, jitter=1e-2, seed=2,
            workload="NSF2(VNNGP(NSF_RBF), K=8) N=4000 M=1000 L=10 G=2000 E=10 (BASELINE.json configs[2])"),
    4: dict(kind="mggp", N_shard=12500, B=8192, M=2048, L=10, G=2000, E=1, n_groups=10, coord_scale=100.0, lengthscale=3.0,
            jitter=0.1, seed=3,
            workload="NSF2(MGGP_SVGP(MGGP_NSF_RBF, 10 groups)) N=100k (12.5k spots resident per GPU) M=2048 L=10 G=2000, minibatch "
                     "8192 spots per GPU per step (BASELINE.json configs[3])"),
    5: dict(kind="hybrid", N_shard=125000, B=16384, M=4096, L=10, T=10, G=2000, E=1, coord_scale=100.0, lengthscale=1.7, jitter=0.1,
            seed=4,
            workload="Hybrid_NSF2(SVGP(NSF_RBF), GaussianPrior T=10) N=1M (125k spots resident per GPU) M=4096 L=10 G=2000, minibatch "
                     "16384 spots per GPU per step (BASELINE.json configs[4])"),
}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tensor=p["bf16_tflops"], tensor_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured")
    except Exception:
        return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi streaming sampler (one long-lived process, 20 ms period) for the clocks line of the timing rules."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.thread, self.n0 = index, None, [], None, 0

    def _pump(self):
        for line in self.proc.stdout:
            if line.strip():
                self.lines.append(line)

    def start(self):
        """Returns once nvidia-smi is streaming (first sample seen, at most 3 s), so that the samples counted from here on
        fall inside the timed region."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 3.0:
                time.sleep(0.01)
            self.n0 = len(self.lines)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            if self.thread is not None:
                self.thread.join(timeout=2)
            rows = [[c.strip() for c in l.split(",")] for l in self.lines[max(0, self.n0 - 1):] if l.strip()]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in rows)]
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_min_mhz=sm[0] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(rows))


# ------------------------------------------------------------------------------------------------
# problems and models
# ------------------------------------------------------------------------------------------------
def make_problem(c, n_spots, seed, dt):
    from gpzoo_b200 import synthetic
    return synthetic.nsf_problem(N=n_spots, M=c["M"], L=c["L"], G=c["G"], E=c["E"], seed=seed, coord_scale=c["coord_scale"],
                                 lengthscale=c["lengthscale"], jitter=c["jitter"], dtype=dt, n_groups=c.get("n_groups", 0))


def build_model(c, prob, dt, dev):
    """The gpzoo_b200 module tree of the config, parameters taken from `prob`.  Returns (model, shared parameters)."""
    import torch
    import gpzoo_b200 as gz
    P = lambda t: torch.nn.Parameter(t.to(dev, dt))
    L, M = prob["mu"].shape
    D = prob["X"].shape[1]
    kind = c["kind"]
    if kind == "mggp":
        kern = gz.kernels.MGGP_NSF_RBF(L=L, n_groups=c["n_groups"])
        kern.set_group_distances(prob["group_distances"].float())
        kern.embedding = torch.nn.Parameter(kern.embedding.to(dev, dt), requires_grad=False)
        kern.group_diff_param = P(prob["gdp"])
        gp = gz.gp.MGGP_SVGP(kern, dim=D, M=M, jitter=prob["jitter"], n_groups=c["n_groups"])
        gp.groupsZ = torch.nn.Parameter(prob["groupsZ"].to(dev), requires_grad=False)
    else:
        kern = gz.kernels.NSF_RBF(L=L)
        if kind == "vnngp":
            gp = gz.gp.VNNGP(kern, dim=D, M=M, K=c["K"], jitter=prob["jitter"])
        else:
            gp = gz.gp.SVGP(kern, dim=D, M=M, jitter=prob["jitter"])
    kern.sigma, kern.lengthscale = P(prob["sigma"]), P(prob["lengthscale"])
    gp.Z, gp.mu, gp.Lu = P(prob["Z"]), P(prob["mu"]), P(prob["Lu_raw"])
    shared = [gp.Z, kern.sigma, kern.lengthscale, gp.mu, gp.Lu]
    if kind == "mggp":
        shared.append(kern.group_diff_param)
    if kind == "hybrid":
        n = prob["V"].shape[0]
        g = torch.Generator().manual_seed(c["seed"] + 77)
        prior = gz.gp.GaussianPrior(prob["y"][:, :1], L=c["T"])
        prior.mean = P(0.1 * torch.randn(c["T"], n, generator=g, dtype=torch.float64))
        prior.scale = P(torch.rand(c["T"], n, generator=g, dtype=torch.float64))
        model = gz.likelihoods.Hybrid_NSF2(gp, prior, prob["y"][:, :1], L=L, T=c["T"])
        model.sf.W, model.V = P(prob["W"]), P(prob["V"])
        model.cf.W = P(torch.rand(c["G"], c["T"], generator=g, dtype=torch.float64))
        shared += [model.sf.W, model.cf.W]
    else:
        model = gz.likelihoods.NSF2(gp, prob["y"][:, :1], L=L)
        model.W, model.V = P(prob["W"]), P(prob["V"])
        shared.append(model.W)
    return model, shared


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU PyTorch path
# ------------------------------------------------------------------------------------------------
def _oracle_step(c, n_sample, dt, device="cpu"):
    """(run, cleanup-free) closure of ONE oracle fwd+bwd step of config 2's model at n_sample spots."""
    import torch
    from oracle import gpzoo_oracle as O
    prob = make_problem(c, n_sample, c["seed"], dt)
    p = O.NSFParams(**{k: prob[k].clone().to(device) for k in ("Z", "sigma", "lengthscale", "mu", "Lu_raw", "W", "V")}, jitter=prob["jitter"])
    X, y, eps = prob["X"].to(device), prob["y"].to(device), prob["eps"].to(device)
    return (lambda: O.value_and_grads(lambda: O.nsf_svgp_terms(p, X, y, eps), p.leaves())), prob


def cpu_reference_full_step(c, steps, budget_s=120.0, sizes=(2048, 4096)):
    """Time the oracle port (the reference's op sequence on CPU torch, fp32, all host threads) at two bounded spot counts with
    the full M / L / G and extrapolate to the full N: the step costs t(N) = a + b N (a: the O(M^3) Cholesky / KL work,
    b: everything per spot).  `steps` timed steps at the smaller size (fewer if `budget_s` runs out) and one at the larger."""
    import torch
    torch.set_num_threads(os.cpu_count())
    n1, n2 = sizes
    run1, _ = _oracle_step(c, n1, torch.float32)
    run1()                                       # warm-up (thread pools, allocator)
    times, t_start = [], time.perf_counter()
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        run1()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    t1 = sorted(times)[len(times) // 2]
    run2, _ = _oracle_step(c, n2, torch.float32)
    t0 = time.perf_counter()
    run2()
    t2 = time.perf_counter() - t0
    b = max(0.0, (t2 - t1) / (n2 - n1))
    a = max(0.0, t1 - b * n1)
    full = a + b * c["N"]
    cores = torch.get_num_threads()
    note = (f"oracle port of the reference CPU path (torch {cores} threads, fp32), M/L/G full: median {t1:.2f} s/step at N={n1} "
            f"({len(times)} timed), {t2:.2f} s at N={n2} (1 timed); affine fit t = {a:.2f} s + {b * 1e3:.3f} ms x N extrapolated to "
            f"N={c['N']} (EXTRAPOLATION: a full-size CPU step takes about a minute)")
    return full, cores, note


def eager_gpu_reference_step(c, steps, warmup):
    """The same oracle port of the reference's op sequence with every tensor on the GPU — i.e. the reference's stock
    PyTorch-eager path (cuBLAS / cuSOLVER / ATen kernels) on this B200 at the FULL config, CUDA-event timed."""
    import torch
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    run, _ = _oracle_step(c, c["N"], torch.float32, dev)
    for _ in range(max(1, warmup)):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, torch.cuda.max_memory_allocated(dev) / 1e9


def common_config(c):
    """The `config` object both arms print (identical, so the driver can tell that they ran the same job)."""
    return dict(workload=c["workload"], dtype="f32", data="synthetic Slide-seq-shaped (gpzoo_b200.synthetic.nsf_problem, seed %d)" % c["seed"],
                inducing="jittered 32x32 grid", step="ELBO forward + backward, all parameter gradients, optimiser excluded")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[2]
    full, cores, note = cpu_reference_full_step(c, args.steps)
    val = 1.0 / full
    line = dict(metric=METRIC, value=val, unit="steps/s", impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=full * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                config=common_config(c),
                cpu_baseline=dict(value=val, unit="steps/s", cores=cores, kind="port", sample=note),
                e2e=dict(value=val, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    try:
        import torch
        if torch.cuda.is_available() and args.ref_device != "cpu":
            ms, gb = eager_gpu_reference_step(c, 3, 1)
            line["eager_gpu"] = dict(ms_per_step=ms, value=1e3 / ms, unit="steps/s", peak_mem_gb=gb,
                                     note="the reference's op sequence (oracle port) as stock PyTorch eager on this B200, fp32, FULL "
                                          "config: the existing-GPU-path bar of SURVEY.md §0")
    except Exception as e:                       # the CPU number stands on its own
        line["eager_gpu"] = dict(error=str(e)[:200])
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from gpzoo_b200 import _cabi, functional
    from gpzoo_b200.distributed import FlatGradReducer, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    functional.set_sync_checks(False)          # Cholesky info / overflow guards are checked once, after the timed region
    # optional: the backward of the replicated O(M^3) chain factor-sharded over the ranks (functional.set_chain_sharding).
    # Measured on 8 B200: 4.08 ms/step against 3.92 ms without — the all-to-all and the repacking cost what the smaller GEMMs
    # save — so it is off unless asked for.
    shard_chain = world > 1 and args.chain_shard
    if shard_chain:
        functional.set_chain_sharding(dist.group.WORLD)
    dt = torch.float32
    c = CONFIGS[args.config]
    kind = c["kind"]
    minibatched = kind in ("mggp", "hybrid")

    def setup(mode):
        """strong: the N spots of the config are split over the ranks (the same job at every world size);
        weak: every rank owns a full-size shard (config 2: 32768 spots each; configs 4/5: the per-GPU resident shard and a
        fixed minibatch per GPU per step, as those configs are defined).  Returns a dict with the model, the shared parameters,
        pinned host copies of the rank's inputs, and the device-resident step inputs."""
        if minibatched:
            prob = make_problem(c, c["N_shard"], c["seed"], dt)
            if rank > 0:                        # same parameters on every rank, a different block of spots
                data = make_problem(c, c["N_shard"], c["seed"] + 1000 + rank, dt)
                for k in ("X", "y", "eps", "V", "groupsX"):
                    if k in data:
                        prob[k] = data[k]
            sl = slice(0, c["N_shard"])
        elif mode == "weak":
            prob = make_problem(c, c["N"], c["seed"], dt)
            if rank > 0:
                data = make_problem(c, c["N"], c["seed"] + 1000 + rank, dt)
                for k in ("X", "y", "eps", "V"):
                    prob[k] = data[k]
            sl = slice(0, c["N"])
        else:
            prob = make_problem(c, c["N"], c["seed"], dt)
            lo, hi = shard_range(c["N"], world, rank)
            sl = slice(lo, hi)
        st = dict(prob=prob, sl=sl)
        st["hX"] = prob["X"][sl].contiguous().pin_memory()
        st["hy"] = prob["y"][:, sl].contiguous().pin_memory()
        ploc = dict(prob)
        ploc["V"] = prob["V"][sl].contiguous()
        st["model"], st["shared"] = build_model(c, ploc, dt, dev)
        st["X"], st["y"] = st["hX"].to(dev), st["hy"].to(dev)
        st["eps"] = prob["eps"][:, :, sl].contiguous().to(dev)
        st["gX"] = prob["groupsX"][sl].to(dev) if "groupsX" in prob else None
        if minibatched:                         # one fixed minibatch (timing does not depend on which spots it holds)
            g = torch.Generator().manual_seed(c["seed"] + 5 + rank)
            st["idx"] = torch.randperm(c["N_shard"], generator=g)[:c["B"]].to(dev)
            st["eps"] = st["eps"][:, :, st["idx"]].contiguous()
        st["reducer"] = FlatGradReducer(st["shared"], device=dev, dtype=dt)
        return st

    DEBUG = os.environ.get("GPZ_BENCH_DEBUG") == "1"

    def step(st, Xd=None, yd=None, eps="given"):
        model = st["model"]
        for p in model.parameters():
            p.grad = None
        Xd = st["X"] if Xd is None else Xd
        yd = st["y"] if yd is None else yd
        e = st["eps"] if eps == "given" else None
        kw = dict(E=c["E"], eps=e, kl_weight=1.0 / world)
        if minibatched:
            kw["idx"] = st["idx"]
            if st["gX"] is not None:
                kw["groupsX"] = st["gX"][st["idx"]]
        if DEBUG:
            _cabi.host_profile = {}
            t0 = time.perf_counter()
        elbo = model.elbo(Xd, yd, **kw)
        if DEBUG:
            t1 = time.perf_counter()
        (-elbo).backward()
        if DEBUG:
            t2 = time.perf_counter()
        out = st["reducer"].all_reduce(elbo)       # one NCCL all-reduce of the flat shared-gradient buffer (+ ELBO)
        if DEBUG:
            ms_ = torch.cuda.memory_stats()
            na = ms_.get("num_device_alloc", 0), ms_.get("num_device_free", 0), ms_.get("num_alloc_retries", 0)
            if na != st.get("_na"):
                print("allocator: cudaMalloc/cudaFree/retries so far", na, "reserved GB %.2f" % (torch.cuda.memory_reserved() / 1e9),
                      file=sys.stderr, flush=True)
                st["_na"] = na
        if DEBUG and time.perf_counter() - t0 > 0.03:
            print("slow step on the host: fwd %.1f ms, bwd %.1f ms, reduce %.1f ms; longest C-ABI calls (ms): %s" % (
                (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3,
                {k: round(v, 1) for k, v in _cabi.host_profile.items() if v > 1.0}), file=sys.stderr, flush=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        per = []
        for _ in range(n):
            t1 = time.perf_counter()
            fn()
            per.append((time.perf_counter() - t1) * 1e3)
        host_ms[0] = (time.perf_counter() - t0) * 1e3 / n      # host time to ENQUEUE one step (no synchronisation inside)
        if os.environ.get("GPZ_BENCH_DEBUG") == "1":
            print("host ms per step:", " ".join(f"{v:.1f}" for v in per), file=sys.stderr, flush=True)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    warm = max(3, args.warmup)
    main_mode = "weak" if minibatched else "strong"
    st = setup(main_mode)
    for _ in range(warm):
        step(st)
    functional.check_cholesky_info()
    # first pass over the K steps with a CUDA-event pair around every C-ABI call: per-call times for the roofline block.  It runs
    # BEFORE the headline pass so that one-off host stalls of a young process (allocator growth, NCCL connection set-up: a 60 ms
    # gap was seen once in the first 20 steps of a 2-rank run with W = 3) land here and not in the headline's 0.1-0.2 s.
    _cabi.profile = {}
    timed(lambda: step(st), args.steps)
    prof, _cabi.profile = _cabi.profile, None
    functional.check_cholesky_info()
    # A full (generation-2) collection of the CPython garbage collector walks every live object of the interpreter -- about
    # 100 ms with torch imported -- and fires at allocation-count-dependent moments; one inside a 170 ms timed region doubles
    # the reading (seen as 19 ms/step with 100 ms host stalls, tools/ + DESIGN.md section 7).  Freezing moves everything
    # allocated so far out of the collector's reach; the garbage of the steps themselves is still collected.
    gc.collect()
    gc.freeze()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    k0 = _cabi.kernel_launches()
    ms = timed(lambda: step(st), args.steps)          # the headline: no per-call instrumentation inside
    launches = (_cabi.kernel_launches() - k0) // args.steps
    host_enqueue_ms = host_ms[0]
    clocks = sampler.stop() if rank == 0 else None
    functional.check_cholesky_info()
    n_loc = int(st["idx"].numel()) if minibatched else int(st["X"].shape[0])

    # multi-GPU numeric check (strong leg): the all-reduced ELBO and shared gradients of the sharded step against the same
    # 32768-spot step computed by ONE rank alone
    dp_parity = None
    if world > 1 and main_mode == "strong":
        total = step(st)
        got = [p.grad.detach().clone() for p in st["shared"]]
        if rank == 0:
            full = dict(st["prob"])
            functional.set_chain_sharding(None)          # the reference step is computed by this rank alone
            m1, sh1 = build_model(c, full, dt, dev)
            e1 = m1.elbo(full["X"].to(dev), full["y"].to(dev), E=c["E"], eps=full["eps"].to(dev))
            (-e1).backward()
            if shard_chain:
                functional.set_chain_sharding(dist.group.WORLD)
            rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
            names = ["Z", "sigma", "lengthscale", "mu", "Lu", "W"]
            errs = {n: rel(g, p.grad) for n, g, p in zip(names, got, sh1)}
            errs["elbo"] = rel(total, e1.detach())
            dp_parity = dict(rel_err_vs_1gpu=errs, tol=1e-4, ok=bool(max(errs.values()) < 1e-4))
            if not dp_parity["ok"]:
                dp_parity["norms"] = {n: (float(g.norm()), float(p.grad.norm())) for n, g, p in zip(names, got, sh1)}
            del m1, sh1, e1
            assert dp_parity["ok"], dp_parity

    # e2e: this rank's inputs start in pinned host memory every step and the loss is read back to the host every step.
    # The upload of step i+1 runs on a copy stream while step i computes (double buffering, the usual input pipeline of a
    # training loop); every upload and every read-back lies inside the timed region: n steps = n uploads + n read-backs.
    ms_e2e, h2d = None, 0
    if not args.no_e2e and not minibatched:     # (configs 4/5 keep their shard resident and index minibatches on the device)
        # The landing buffers below are written on the copy stream; the allocator hands them out in main-stream order, so they
        # may be memory that still-queued main-stream work (a rank that runs ahead of its peers has whole steps queued behind an
        # all-reduce) has not finished with.  Drain everything first.
        barrier()
        copy_stream = torch.cuda.Stream(device=dev)
        hX, hy = st["hX"], st["hy"]
        dbuf = [(torch.empty_like(st["X"]), torch.empty_like(st["y"])) for _ in range(2)]
        free_ev = [None, None]                                                      # "the step that read buffer k has finished"
        h2d = hX.numel() * 4 + hy.numel() * 4

        def upload(k):
            with torch.cuda.stream(copy_stream):
                if free_ev[k] is not None:
                    copy_stream.wait_event(free_ev[k])
                dbuf[k][0].copy_(hX, non_blocking=True)
                dbuf[k][1].copy_(hy, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return dbuf[k][0], dbuf[k][1], ev

        host_loss = [torch.empty((), dtype=dt).pin_memory() for _ in range(2)]

        def e2e_steps(n):
            """n steps, n uploads, n loss read-backs.  The loss of step i is copied to pinned host memory asynchronously and
            read on the host after step i+1 has been queued (one-step lag, the usual logging pattern), so neither the upload
            nor the read-back leaves the GPU idle; the last read-back is waited for before the timed region ends."""
            cur = torch.cuda.current_stream()
            nxt = upload(0)
            pending = None
            losses = []
            for i in range(n):
                Xd, yd, ev = nxt
                cur.wait_event(ev)
                if i + 1 < n:
                    nxt = upload((i + 1) & 1)
                loss = step(st, Xd, yd, eps=None)              # eps drawn on the device
                host_loss[i & 1].copy_(loss.detach().reshape(()), non_blocking=True)      # loss D2H
                done = torch.cuda.Event()
                done.record(cur)
                free_ev[i & 1] = done
                if pending is not None:
                    pending[0].synchronize()
                    losses.append(float(pending[1]))
                pending = (done, host_loss[i & 1])
            pending[0].synchronize()
            losses.append(float(pending[1]))
            assert len(losses) == n and all(v == v for v in losses)

        e2e_steps(2)
        n_e2e = max(2, args.steps)                   # K steps like the device-resident leg (the first upload is exposed: pipeline fill)
        ms_e2e = timed(lambda: e2e_steps(n_e2e), 1) / n_e2e
        del dbuf

    # second leg (config 2, N > 1): weak scaling, every GPU runs the full 32768 spots
    weak = None
    if world > 1 and main_mode == "strong" and not args.no_weak:
        del st
        torch.cuda.empty_cache()
        st = setup("weak")
        for _ in range(3):
            step(st)
        ms_w = timed(lambda: step(st), args.steps)
        weak = dict(value=world * 1e3 / ms_w, unit="32768-spot steps/s summed over the GPUs", ms_per_step=ms_w,
                    spots_per_gpu=int(st["X"].shape[0]), global_spots=int(st["X"].shape[0]) * world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    timed_s = ms * args.steps / 1e3
    tensor_peak, tensor_peak_name = (pk["tensor_sustained"], "sustained") if timed_s >= 2.0 else (pk["tensor"], "burst")
    N, M, L, G, E = n_loc, c["M"], c["L"], c["G"] + 0, c["E"]
    F = L + c.get("T", 0)
    algo = {   # per-call algorithmic work on ONE rank: ("hbm", bytes) or ("tensor", flops)
        "svgp_predict_fwd_h": ("tensor", 2.0 * L * M * M * N),
        # merged backward (SvgpMomentsH): gA, S1 = tril(AW A^T), gKzx, each one triangular M x M x N product; the two-node path
        # has four (gA, gKzx, gT, gLinv)
        "svgp_predict_bwd_h": ("tensor", (3.0 if "svgp_chain_bwd_s1" in prof else 4.0) * L * M * M * N),
        "kernel_build_fwd_h": ("hbm", 4.0 * L * M * N),    # Kzx written once as two fp16 planes
        "kernel_build_bwd": ("hbm", 4.0 * L * M * N),      # dL/dKzx read once (the largest of the two calls: the Kzx one)
        "poisson_fwdbwd": ("hbm", 4.0 * G * N + 4.0 * (3 * E * F * N + 2 * G * F + 2 * N)),
        "vnngp_fwd": ("hbm", 8.0 * N * c.get("K", 0) + 4.0 * L * N * (2 * c.get("K", 0) ** 2 + 2 * c.get("K", 0)) + 8.0 * L * N),
    }
    per_call = {}
    for name, evs in prof.items():
        tot = sum(a.elapsed_time(b) for a, b in evs)
        per_call[name] = dict(ms_per_step=tot / args.steps, calls_per_step=len(evs) / args.steps)
    kernels = []
    for name, (bound, work) in algo.items():
        if name not in per_call or work <= 0:
            continue
        evs = prof[name]
        dur = sorted((a.elapsed_time(b) for a, b in evs), reverse=True)[:args.steps]      # the largest call of each step
        avg = sum(dur) / len(dur)
        if bound == "hbm":
            ach, peak, unit = work / avg / 1e6, pk["hbm"], "GB/s"
        else:
            ach, peak, unit = work / avg / 1e9, tensor_peak, "TFLOP/s"
        k = dict(kernel=name, bound=bound, ms=avg, achieved=ach, peak=peak, unit=unit, frac=ach / peak)
        if k["frac"] > 1.2:                    # a fraction far above 1 means the timed call is not doing the credited work
            k["invalid"] = "frac > 1.2: not evidence"
        kernels.append(k)
    dom = max(kernels, key=lambda k: k["ms"]) if kernels else None
    roofline = None
    # DRAM traffic per call (dram__bytes_read.sum + dram__bytes_write.sum over the call's launches) from the committed
    # `ncu --set full` capture of THIS round (profiles/r2_ncu_full_summary.txt); only valid for the workload it was captured on
    NCU_TRAFFIC = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"))) if os.path.exists(
        os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) else {}
    if dom:
        traffic = NCU_TRAFFIC.get(dom["kernel"]) if (args.config, N, world) == (2, 32768, 1) else None
        roofline = dict(kernel=dom["kernel"], bound=dom["bound"], achieved=dom["achieved"], peak=dom["peak"], unit=dom["unit"],
                        frac=dom["frac"], traffic=traffic, traffic_unit="B/call (ncu, profiles/r2_ncu_full_summary.txt)",
                        peak_source=pk["src"] + (f" (bf16 dense, {tensor_peak_name}: the timed region is {timed_s:.2f} s)"
                                                 if dom["bound"] == "tensor" else ""),
                        note="split-FP16 arithmetic issues 3 f16 MMAs per product: a perfect kernel reads frac = 0.333")
    spots = n_loc * world
    cfg = common_config(c)
    line = dict(metric=METRIC if args.config == 2 else METRIC + f" (config {args.config})",
                value=(1e3 / ms) if main_mode == "strong" else world * 1e3 / ms, unit="steps/s", n_gpus=world, steps=args.steps,
                warmup=warm, ms_per_step=ms, higher_is_better=True, scaling=main_mode, vs_baseline=None, dtype="f32", data="synthetic",
                config=cfg,
                run=dict(spots_per_gpu=n_loc, global_spots_per_step=spots, spots_per_s=spots * 1e3 / ms,
                         parallelism=f"dp{world} over spots, 1 NCCL all-reduce/step" + (
                             "; backward of the replicated O(M^3) chain sharded by factor (1 all-to-all)" if shard_chain else ""),
                         l2="inputs larger than L2 (y %.0f MB, Kzx planes %.2f GB per GPU)" % (4e-6 * G * n_loc, 4e-9 * L * M * n_loc),
                         e2e="per step: X, y of the rank's spots uploaded from pinned host memory on a copy stream (double buffered) + "
                             "loss read back to the host with one-step lag",
                         order="W warm-up steps, K steps with per-call CUDA events (per_call_ms / kernels), then the K timed steps of "
                               "the headline, then the K e2e steps"),
                clocks=clocks, gpu_launches=int(launches), host_enqueue_ms_per_step=host_enqueue_ms,
                e2e=(dict(value=((1e3 / ms_e2e) if main_mode == "strong" else world * 1e3 / ms_e2e), unit="steps/s",
                          h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=4) if ms_e2e else None),
                roofline=roofline, kernels=kernels, per_call_ms=per_call)
    if weak is not None:
        line["weak"] = weak
    if dp_parity is not None:
        line["dp_parity"] = dp_parity
    if world == 1 and args.config == 2 and not args.no_cpu_baseline:
        full, cores, note = cpu_reference_full_step(c, 1)
        line["cpu_baseline"] = dict(value=1.0 / full, unit="steps/s", cores=cores, kind="port", sample=note)
        line["parity"] = parity_block(c, dt, dev)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def parity_block(c, dt, dev, n_spots=768):
    """The benchmarked arithmetic against the oracle on the SAME inputs: config 2's full M / L / G at `n_spots` spots; the fp32
    tcgen05 step on the GPU vs the oracle in fp64 on the host (relative L2 per tensor), with the oracle's own fp32 beside it."""
    import torch
    from gpzoo_b200 import _cabi, functional
    from oracle import gpzoo_oracle as O
    prob = make_problem(c, n_spots, c["seed"], torch.float64)
    rel = lambda a, b: float((a.detach().double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-300))

    def oracle(dtype):
        p = O.NSFParams(**{k: prob[k].to(dtype).clone() for k in ("Z", "sigma", "lengthscale", "mu", "Lu_raw", "W", "V")}, jitter=prob["jitter"])
        out, g = O.value_and_grads(lambda: O.nsf_svgp_terms(p, prob["X"].to(dtype), prob["y"].to(dtype), prob["eps"].to(dtype)), p.leaves())
        return dict(elbo=out["elbo"], mean=out["mean"], var=out["var"], dZ=g["Z"], dsigma=g["sigma"], dlengthscale=g["lengthscale"],
                    dmu=g["mu"], dLu=g["Lu_raw"], dW=g["W"], dV=g["V"])
    truth, ref32 = oracle(torch.float64), oracle(torch.float32)
    model, _ = build_model(c, prob, dt, dev)
    _cabi.profile = {}
    elbo, parts = model.elbo(prob["X"].to(dev, dt), prob["y"].to(dev, dt), E=c["E"], eps=prob["eps"].to(dev, dt), return_parts=True)
    elbo.backward()
    calls, _cabi.profile = sorted(_cabi.profile), None
    functional.check_cholesky_info()
    gp = model.prior
    ours = dict(elbo=elbo, mean=parts["mean"], var=parts["var"].clamp(min=1e-6), dZ=gp.Z.grad, dsigma=gp.kernel.sigma.grad,
                dlengthscale=gp.kernel.lengthscale.grad, dmu=gp.mu.grad, dLu=gp.Lu.grad, dW=model.W.grad, dV=model.V.grad)
    errs = {k: rel(ours[k], truth[k]) for k in truth}
    floor = {k: rel(ref32[k], truth[k]) for k in truth}
    return dict(spots=n_spots, against="oracle fp64 on the host, same inputs (M, L, G, jitter, coordinates of configs[1])",
                rel_l2=errs, max_rel_l2=max(errs.values()), tol=1e-4, ok=bool(max(errs.values()) < 1e-4),
                reference_fp32_rel_l2=floor, c_abi_calls=calls)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json config (1-based): 2 = configs[1] (the metric's), 3 VNNGP, 4 MGGP M=2048, 5 hybrid M=4096")
    ap.add_argument("--ref-device", default="auto", choices=["auto", "cpu"],
                    help="--impl reference: 'cpu' skips the PyTorch-eager-on-GPU context number")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling leg")
    ap.add_argument("--chain-shard", action="store_true", help="N > 1: shard the O(M^3) chain's backward by factor over the ranks (A/B)")
    ap.add_argument("--N", type=int, default=None, help="override the number of spots of config 2 (debugging only)")
    args = ap.parse_args()
    if args.N:
        CONFIGS[2]["N"] = args.N
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
