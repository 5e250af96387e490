"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported read-only from
/root/reference/code) in the build container.  The reference cannot travel to the GPU box,
so its outputs are committed as small fixtures together with this script.

    python oracle/make_golden.py            # rewrites tests/golden/

Everything random is derived from integer seeds with CPU generators, so the fixtures hold
only outputs (plus a few small inputs); weights come from `fidm_b200.utils.synth`.
While generating, the oracle restatement (oracle/*.py) is checked against the reference
and the script aborts if they disagree -- this is what "pins" the oracle.
"""
import json
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/code")

import numpy as np  # noqa: E402
import torch  # noqa: E402

import fidm_b200  # noqa: E402,F401
from fidm_b200.arch import CONFIGS, param_shapes, unet_topology  # noqa: E402
from fidm_b200.utils.synth import synth_batch, synth_state_dict  # noqa: E402
from oracle import diffusion_oracle as dor  # noqa: E402
from oracle import unet_oracle as uor  # noqa: E402

import gaussian_diffusion as ref_gd  # noqa: E402  (reference)
from unet import DiffusionInpaintingModel, UNetModel  # noqa: E402  (reference)
from utils.schedules import create_gaussian_diffusion, get_named_beta_schedule  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(os.cpu_count())


_REAL_RANDN = torch.randn


def seeded_noise(kind, t, shape, seed):
    g = torch.Generator().manual_seed(seed * 100003 + {"xT": 0, "inject": 1, "step": 2}[kind] * 50021 + int(t))
    return _REAL_RANDN(*shape, generator=g)


class PatchedRandn:
    """Route the reference's torch.randn / randn_like draws to seeded_noise, in its known call
    order: randn(shape) once, then per step [randn_like(gt) on cache miss], randn_like(x)."""

    def __init__(self, T, seed, inject=True):
        self.seq = [("xT", 0)]
        for t in range(T - 1, -1, -1):
            if inject:
                self.seq.append(("inject", t))
            self.seq.append(("step", t))
        self.i, self.seed = 0, seed

    def _next(self, shape):
        kind, t = self.seq[self.i]
        self.i += 1
        return seeded_noise(kind, t, tuple(shape), self.seed)

    def __enter__(self):
        self._r, self._rl = torch.randn, torch.randn_like
        torch.randn = lambda *s, **k: self._next(s[0] if len(s) == 1 and not isinstance(s[0], int) else s)
        torch.randn_like = lambda x, **k: self._next(x.shape)
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._r, self._rl


def build_ref_model(name, seed):
    cfg = CONFIGS[name]
    c3 = dict(cfg, in_channels=3)
    model = DiffusionInpaintingModel(UNetModel(**c3), in_channels=9).eval()
    sd = synth_state_dict(cfg, seed=seed)
    model.load_state_dict(sd, strict=True)
    return model, sd, cfg


def golden_topology():
    info = {}
    for name, cfg in CONFIGS.items():
        c3 = dict(cfg, in_channels=3)
        with torch.device("meta"):
            m = DiffusionInpaintingModel(UNetModel(**c3))
        info[name] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        mine = [["base_model." + k, list(s)] for k, s in param_shapes(unet_topology(**cfg))]
        assert mine == info[name], name
    with open(os.path.join(OUT, "state_dict_layout.json"), "w") as f:
        json.dump(info, f)


def golden_schedules():
    out = {}
    for sched in ("linear", "cosine", "quadratic", "sqrt"):
        for T in (50, 100, 1000):
            d = create_gaussian_diffusion(steps=T, learn_sigma=True, noise_schedule=sched)
            tab = dor.Tables(get_named_beta_schedule(sched, T))
            for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                      "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                      "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                      "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"):
                assert np.array_equal(getattr(d, k), getattr(tab, k)), (sched, T, k)
            idx = [0, 1, T // 2, T - 1]
            out[f"{sched}_{T}"] = {k: [float(getattr(d, k)[i]) for i in idx]
                                   for k in ("betas", "alphas_cumprod", "posterior_log_variance_clipped",
                                             "posterior_mean_coef1", "sqrt_recipm1_alphas_cumprod")}
    with open(os.path.join(OUT, "schedules.json"), "w") as f:
        json.dump(out, f)


def golden_steps():
    """Single reverse steps of the reference on tiny tensors, every mode K4 implements."""
    B, C, H, W = 2, 3, 8, 8
    cases = []
    seed = 11
    for sched, T in (("cosine", 100), ("linear", 1000), ("quadratic", 100)):
        for learn_sigma, sigma_small in ((True, False), (False, False), (False, True)):
            d = create_gaussian_diffusion(steps=T, learn_sigma=learn_sigma, sigma_small=sigma_small,
                                          noise_schedule=sched)
            tab = dor.Tables(get_named_beta_schedule(sched, T))
            vt = "learned_range" if learn_sigma else ("fixed_small" if sigma_small else "fixed_large")
            for t in (0, 1, T // 2, T - 1):
                for mode, eta, cumulative in (("ddim", 0.0, True), ("ddim", 0.7, False), ("ddpm", 0.0, True)):
                    seed += 1
                    g = torch.Generator().manual_seed(seed)
                    x = torch.randn(B, C, H, W, generator=g) * (1.0 + t / T)
                    gt = torch.rand(B, C, H, W, generator=g) * 2 - 1
                    keep = (torch.rand(B, 1, H, W, generator=g) > 0.4).float()
                    mo = torch.randn(B, 2 * C if learn_sigma else C, H, W, generator=g)
                    n_inj = torch.randn(B, C, H, W, generator=g)
                    z = torch.randn(B, C, H, W, generator=g)
                    tt = torch.full((B,), t, dtype=torch.int64)
                    d.clear_gt_noise_cache()
                    draws = iter([n_inj, z])
                    rl = torch.randn_like
                    torch.randn_like = lambda a, **k: next(draws)
                    try:
                        fn = d.ddim_sample if mode == "ddim" else d.p_sample
                        kw = dict(eta=eta) if mode == "ddim" else {}
                        got = fn(lambda xx, ts, **k: mo, x, tt, clip_denoised=True,
                                 model_kwargs={"gt": gt, "gt_keep_mask": keep},
                                 use_inpainting_injection=True, injection_schedule="all",
                                 use_cumulative_noise=cumulative, **kw)
                        d.clear_gt_noise_cache()
                        draws = iter([n_inj])
                        x_inj = d.apply_inpainting_injection(x, tt, gt, keep, use_cumulative_noise=cumulative)
                    finally:
                        torch.randn_like = rl
                    # pin the oracle: bit-exact on pure mul/add paths, 1e-6 when exp() is involved
                    xi = dor.inject(tab, x, t, gt, keep, n_inj, cumulative)
                    assert torch.equal(xi, x_inj)
                    if mode == "ddim":
                        s, x0 = dor.ddim_update(tab, mo, xi, t, z, eta, vt)
                        assert torch.equal(s, got["sample"]) and torch.equal(x0, got["pred_xstart"])
                    else:
                        s, x0 = dor.ddpm_update(tab, mo, xi, t, z, vt)
                        assert torch.equal(s, got["sample"]) and torch.equal(x0, got["pred_xstart"])
                    cases.append(dict(sched=sched, T=T, var_type=vt, t=t, mode=mode, eta=eta,
                                      cumulative=cumulative, seed=seed,
                                      x_inj=x_inj.clone(), sample=got["sample"].clone(),
                                      pred_xstart=got["pred_xstart"].clone()))
    torch.save(cases, os.path.join(OUT, "sampler_steps.pt"))
    print("sampler step cases:", len(cases))


def golden_t64():
    model, sd, cfg = build_ref_model("T64", seed=1)
    B = 2
    data = synth_batch(B, 64, seed=3)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 3, 64, 64, generator=g)
    t = torch.tensor([37, 5], dtype=torch.int64)
    with torch.no_grad():
        out = model(x, t, masked_image=data["masked_image"], mask=data["mask"])
        mine = uor.inpaint_forward(sd, cfg, x, t, data["masked_image"], data["mask"])
    rel = ((out - mine).norm() / out.norm()).item()
    print("T64 forward: |out|", out.norm().item(), "oracle rel-L2", rel)
    assert rel < 2e-6, rel
    torch.save({"seed_weights": 1, "seed_data": 3, "x": x, "t": t, "out": out},
               os.path.join(OUT, "t64_forward.pt"))

    # config #1: T64, DDIM-50 cosine, B=1, injection on (class path), explicit seeded noise
    T, seed = 50, 9
    d = create_gaussian_diffusion(steps=T, learn_sigma=True, noise_schedule="cosine")
    data1 = synth_batch(1, 64, seed=4)
    gt, keep = data1["gt"], data1["gt_keep_mask"]

    def model_fn(xx, ts, gt=None, gt_keep_mask=None, **kw):
        return model(xx, ts, masked_image=gt * gt_keep_mask, mask=1 - gt_keep_mask)

    trace = []
    with PatchedRandn(T, seed), torch.no_grad():
        for o in d.ddim_sample_loop_progressive(model_fn, (1, 3, 64, 64), model_kwargs={"gt": gt, "gt_keep_mask": keep},
                                                device="cpu", eta=0.0, use_inpainting_injection=True):
            trace.append(o)
    final = trace[-1]["sample"]

    tab = dor.Tables(get_named_beta_schedule("cosine", T))
    otrace = []
    with torch.no_grad():
        mine = dor.sample_loop(
            tab, lambda xx, ts, **k: uor.inpaint_forward(sd, cfg, xx, ts, gt * keep, 1 - keep),
            (1, 3, 64, 64), ddim=True, x_T=seeded_noise("xT", 0, (1, 3, 64, 64), seed), gt=gt, keep=keep,
            noise_fn=lambda kind, t: seeded_noise(kind, t, (1, 3, 64, 64), seed), trace=otrace)
    mse = ((final - mine) ** 2).mean().item()
    psnr = 10 * np.log10(4.0 / max(mse, 1e-20))
    print("T64 DDIM-50: oracle vs reference PSNR", psnr)
    assert psnr > 80, psnr
    torch.save({"seed_weights": 1, "seed_data": 4, "seed_noise": seed, "T": T, "final": final,
                "pred_xstart_t25": trace[24]["pred_xstart"], "sample_t25": trace[24]["sample"]},
               os.path.join(OUT, "t64_ddim50.pt"))

    # DDPM (p_sample_loop) on a short linear schedule: exercises learned-range variance + exp
    T2, seed2 = 100, 21
    d2 = create_gaussian_diffusion(steps=T2, learn_sigma=True, noise_schedule="linear")
    with PatchedRandn(T2, seed2), torch.no_grad():
        fin2 = d2.p_sample_loop(model_fn, (1, 3, 64, 64), model_kwargs={"gt": gt, "gt_keep_mask": keep},
                                device="cpu", use_inpainting_injection=True)
    tab2 = dor.Tables(get_named_beta_schedule("linear", T2))
    with torch.no_grad():
        mine2 = dor.sample_loop(
            tab2, lambda xx, ts, **k: uor.inpaint_forward(sd, cfg, xx, ts, gt * keep, 1 - keep),
            (1, 3, 64, 64), ddim=False, x_T=seeded_noise("xT", 0, (1, 3, 64, 64), seed2), gt=gt, keep=keep,
            noise_fn=lambda kind, t: seeded_noise(kind, t, (1, 3, 64, 64), seed2))
    psnr2 = 10 * np.log10(4.0 / max(((fin2 - mine2) ** 2).mean().item(), 1e-20))
    print("T64 DDPM-100: oracle vs reference PSNR", psnr2)
    assert psnr2 > 80, psnr2
    torch.save({"seed_weights": 1, "seed_data": 4, "seed_noise": seed2, "T": T2, "final": fin2},
               os.path.join(OUT, "t64_ddpm100.pt"))


def golden_variants():
    """Non-default ctor switches (additive emb, conv resample, num_heads path)."""
    res = {}
    for tag, over in (("plain", dict(use_scale_shift_norm=False, resblock_updown=False, conv_resample=True)),
                      ("pool", dict(use_scale_shift_norm=True, resblock_updown=False, conv_resample=False))):
        cfg = dict(CONFIGS["T64"], **over, image_size=32, num_heads_upsample=2)
        c3 = dict(cfg, in_channels=3)
        model = DiffusionInpaintingModel(UNetModel(**c3), in_channels=9).eval()
        sd = synth_state_dict(cfg, seed=2)
        model.load_state_dict(sd, strict=True)
        data = synth_batch(2, 32, seed=6)
        g = torch.Generator().manual_seed(8)
        x = torch.randn(2, 3, 32, 32, generator=g)
        t = torch.tensor([3, 44], dtype=torch.int64)
        with torch.no_grad():
            out = model(x, t, masked_image=data["masked_image"], mask=data["mask"])
            mine = uor.inpaint_forward(sd, cfg, x, t, data["masked_image"], data["mask"])
        rel = ((out - mine).norm() / out.norm()).item()
        print("variant", tag, "oracle rel-L2", rel)
        assert rel < 2e-6
        res[tag] = {"cfg": cfg, "seed_weights": 2, "seed_data": 6, "x": x, "t": t, "out": out}
    torch.save(res, os.path.join(OUT, "t32_variants.pt"))


def _reference_script_methods():
    """Method bodies of the reference's InpaintingSampler, extracted from the script source (the script
    imports lpips / skimage / pytorch_fid at module level and cannot be imported here)."""
    import ast
    import textwrap
    import types
    from tqdm.auto import tqdm
    path = "/root/reference/code/test_inp_ddim_100.py"
    src = open(path).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "InpaintingSampler"][0]
    ns = {"torch": torch, "np": np, "tqdm": tqdm}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("model_fn", "create_ddim_timestep_sequence",
                                                           "inpainting_ddim_sample_loop", "inpainting_p_sample_loop"):
            exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns, types


class SeqRandn:
    """torch.randn / randn_like patched to an explicit (kind, t) sequence."""

    def __init__(self, seq, seed):
        self.seq, self.i, self.seed = list(seq), 0, seed

    def _next(self, shape):
        kind, t = self.seq[self.i]
        self.i += 1
        return seeded_noise(kind, t, tuple(shape), self.seed)

    def __enter__(self):
        self._r, self._rl = torch.randn, torch.randn_like
        torch.randn = lambda *s, **k: self._next(s[0] if len(s) == 1 and not isinstance(s[0], int) else s)
        torch.randn_like = lambda x, **k: self._next(x.shape)
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._r, self._rl
        assert self.i == len(self.seq), (self.i, len(self.seq))


def golden_script_loops():
    """The evaluation scripts' own loops (strided DDIM, post-step injection) on T64."""
    import contextlib
    import io
    import types as _t
    from oracle import script_oracle as sor
    ns, types = _reference_script_methods()
    model, sd, cfg = build_ref_model("T64", seed=1)
    data = synth_batch(1, 64, seed=4)
    gt, masks = data["gt"], data["mask"]
    shape = (1, 3, 64, 64)
    out = {}
    for tag, steps, sched, n_ddim, eta in (("quad1000_ddim20", 1000, "quadratic", 20, 0.0),
                                           ("cos100_ddim10_eta", 100, "cosine", 10, 0.5)):
        d = create_gaussian_diffusion(steps=steps, learn_sigma=True, noise_schedule=sched)
        me = _t.SimpleNamespace(model=model, diffusion=d, args=_t.SimpleNamespace(ddim_timesteps=n_ddim))
        for name in ("model_fn", "create_ddim_timestep_sequence", "inpainting_ddim_sample_loop", "inpainting_p_sample_loop"):
            setattr(me, name, types.MethodType(ns[name], me))
        seq = me.create_ddim_timestep_sequence(steps, n_ddim)
        assert list(seq) == list(sor.ddim_timestep_sequence(steps, n_ddim))
        order = [("xT", 0)]
        for t in seq:
            if t > 0 and eta > 0:
                order.append(("step", int(t)))
            if t > 0:
                order.append(("inject", int(t)))
        seed = 31
        with SeqRandn(order, seed), torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            ref = me.inpainting_ddim_sample_loop(me.model_fn, shape, gt, masks, clip_denoised=True, device="cpu", eta=eta)
        tab = dor.Tables(get_named_beta_schedule(sched, steps))
        with torch.no_grad():
            mine = sor.script_ddim_loop(
                tab, lambda x, t, gt=None, gt_keep_mask=None: uor.inpaint_forward(sd, cfg, x, t, gt * gt_keep_mask, 1 - gt_keep_mask),
                shape, gt, masks, n_ddim, eta=eta, x_T=seeded_noise("xT", 0, shape, seed),
                noise_fn=lambda kind, t: seeded_noise(kind, t, shape, seed))
        assert torch.equal(ref, mine), tag
        print("script DDIM", tag, "evals", len(seq), "oracle == reference (bit-exact)")
        out[tag] = {"steps": steps, "sched": sched, "n_ddim": n_ddim, "eta": eta, "seed_noise": seed,
                    "seed_weights": 1, "seed_data": 4, "n_evals": len(seq), "final": ref}
    # DDPM script loop on a short schedule
    steps, seed = 30, 37
    d = create_gaussian_diffusion(steps=steps, learn_sigma=True, noise_schedule="cosine")
    me = _t.SimpleNamespace(model=model, diffusion=d, args=_t.SimpleNamespace(ddim_timesteps=10))
    for name in ("model_fn", "inpainting_p_sample_loop"):
        setattr(me, name, types.MethodType(ns[name], me))
    order = [("xT", 0)]
    for i in range(steps - 1, -1, -1):
        order.append(("step", i))
        if i > 0:
            order.append(("inject", i))
    with SeqRandn(order, seed), torch.no_grad():
        ref = me.inpainting_p_sample_loop(me.model_fn, shape, gt, masks, clip_denoised=True, device="cpu")
    tab = dor.Tables(get_named_beta_schedule("cosine", steps))
    with torch.no_grad():
        mine = sor.script_ddpm_loop(
            tab, lambda x, t, gt=None, gt_keep_mask=None: uor.inpaint_forward(sd, cfg, x, t, gt * gt_keep_mask, 1 - gt_keep_mask),
            shape, gt, masks, x_T=seeded_noise("xT", 0, shape, seed), noise_fn=lambda kind, t: seeded_noise(kind, t, shape, seed))
    assert torch.equal(ref, mine)
    print("script DDPM-30: oracle == reference (bit-exact)")
    out["cos30_ddpm"] = {"steps": steps, "sched": "cosine", "seed_noise": seed, "seed_weights": 1, "seed_data": 4, "final": ref}
    torch.save(out, os.path.join(OUT, "t64_script_loops.pt"))


if __name__ == "__main__":
    golden_topology()
    golden_schedules()
    golden_steps()
    golden_t64()
    golden_variants()
    golden_script_loops()
    print("golden fixtures written to", OUT)
