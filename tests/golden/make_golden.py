"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the fixtures it writes are committed
and are the only thing the tests read.  Usage:  python tests/golden/make_golden.py

What is recorded: for each FG SIREN variant one tiny forward of the reference's
``ImplicitGenerator3d`` with torch.rand/torch.randn replayed from recorded draws, plus
function-level vectors for ``fancy_integration`` and ``sample_pdf`` (including degenerate
inputs).  Stage taps that the reference does not return are captured by wrapping the reference's
own functions (no reference source is modified or copied).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CNG_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

# the reference imports matplotlib.pyplot without using it (volumetric_rendering.py:12)
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)
import generators.generators as ref_gen  # noqa: E402
import generators.volumetric_rendering as ref_vr  # noqa: E402
import generators.siren as ref_siren  # noqa: E402

from oracle import nerf_path as oracle  # noqa: E402

torch.set_num_threads(1)


class Replay:
    """Replace torch.rand / torch.randn by a queue of pre-drawn tensors (order checked)."""

    def __init__(self, queue):
        self.queue = list(queue)

    def _pop(self, kind, shape):
        k, t = self.queue.pop(0)
        assert k == kind and tuple(t.shape) == tuple(shape), (k, kind, t.shape, shape)
        return t.clone()

    def rand(self, *shape, device=None, **kw):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        return self._pop("rand", shape)

    def randn(self, *shape, device=None, **kw):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        return self._pop("randn", shape)


def run_reference_forward(siren_type, state, z, cam, draws, meta):
    ref_name = oracle.resolve_siren_type(siren_type)
    vol0 = z[0] if isinstance(z, tuple) else z
    z_dim = vol0.shape[1] if ref_name in ("TALLSIREN_dRes", "TALLSIREN_dResLong") else 256      # those classes set input_dim = z_dim (siren.py:355, :433)
    gen = ref_gen.ImplicitGenerator3d(ref_name, z_dim=z_dim, input_dim=vol0.shape[1], output_dim=4, hidden_dim=256)
    gen.load_state_dict(state, strict=True)
    gen.set_device(torch.device("cpu"))
    gen.eval()
    q = [("rand", draws["u_jitter"]), ("randn", draws["noise_coarse"])]
    if meta["hierarchical_sample"]:
        q += [("rand", draws["u_resample"]), ("randn", draws["noise_final"])]
    rp = Replay(q)
    taps = {}
    orig = dict(rand=torch.rand, randn=torch.randn, siren_fwd=gen.siren.forward,
                fancy=ref_gen.fancy_integration, spdf=ref_gen.sample_pdf)

    siren_calls = []

    def siren_tap(points, zz, img_size, num_steps):
        out = orig["siren_fwd"](points, zz, img_size, num_steps)
        siren_calls.append((points.detach().clone(), out.detach().clone()))
        return out

    fancy_calls = []

    def fancy_tap(*a, **k):
        out = orig["fancy"](*a, **k)
        fancy_calls.append(tuple(o.detach().clone() for o in out))
        return out

    spdf_calls = []

    def spdf_tap(bins, weights, n, det=False, eps=1e-5):
        out = orig["spdf"](bins, weights, n, det=det, eps=eps)
        spdf_calls.append((bins.clone(), weights.clone(), out.clone()))
        return out

    torch.rand, torch.randn = rp.rand, rp.randn
    gen.siren.forward = siren_tap
    ref_gen.fancy_integration, ref_gen.sample_pdf = fancy_tap, spdf_tap
    try:
        with torch.no_grad():
            pixels, depth = gen(z, cam, **meta)
    finally:
        torch.rand, torch.randn = orig["rand"], orig["randn"]
        ref_gen.fancy_integration, ref_gen.sample_pdf = orig["fancy"], orig["spdf"]
    assert not rp.queue, "reference consumed fewer draws than recorded"
    B, R, S = cam.shape[0], meta["img_size"] ** 2, meta["num_steps"]
    taps["pixels"], taps["depth"] = pixels, depth
    taps["points_coarse"] = siren_calls[0][0].reshape(B, R, S, 3)
    taps["rgb_sigma_coarse"] = siren_calls[0][1].reshape(B, R, S, 4)
    if meta["hierarchical_sample"]:
        taps["weights_coarse"] = fancy_calls[0][2]
        taps["t_fine"] = spdf_calls[0][2].reshape(B, R, S, 1)
        taps["points_fine"] = siren_calls[1][0].reshape(B, R, S, 3)
        taps["rgb_sigma_fine"] = siren_calls[1][1].reshape(B, R, S, 4)
    taps["rgb"], taps["dist"], taps["weights_final"] = fancy_calls[-1]
    return taps


def make_forward_fixture(siren_type, seed, hierarchical=True, clamp_mode="relu", nerf_noise=0.0,
                         white_back=True, last_back=False, img_size=12, S=8, V=12, B=2, feat_std=0.3, dense=False):
    """``dense``: the head rows of the random-init network are scaled by oracle.DENSE_HEAD_GAINS (alpha spans 0..~0.9, rays
    saturate, the image has texture) before the REFERENCE renders it; the gains are stored in the fixture's meta."""
    g = torch.Generator().manual_seed(seed)
    state = oracle.init_generator_state(siren_type, z_dim=256, input_dim=32, hidden_dim=256, seed=seed)
    gains = None
    if dense:
        gains = list(oracle.DENSE_HEAD_GAINS[oracle.resolve_siren_type(siren_type)])
        state = oracle.dense_head_state(state, *gains)
    vol = torch.randn((B, 32, V, V, V), generator=g) * feat_std
    glob = torch.randn((B, 256), generator=g) * 0.05 + 0.19
    rng = np.random.RandomState(seed)
    cam = oracle.look_at_cam2world(oracle.random_camera_origins(B, 0.7, 1.5, "y", rng), "y")
    # also pin the camera helpers against the reference
    rng2 = np.random.RandomState(seed)
    st = np.random.get_state()
    np.random.set_state(rng2.get_state())
    o_ref = ref_vr.sample_camera_positions(torch.device("cpu"), "y", 0.7, 1.5, B)
    np.random.set_state(st)
    cam_ref = ref_vr.create_cam2world_matrix(o_ref, "y")
    assert torch.equal(cam, cam_ref), "camera helpers diverge from the reference"
    draws = oracle.draw_randoms(B, img_size, S, hierarchical, g)
    meta = dict(img_size=img_size, fov=49.134342641202636, ray_start=0.25, ray_end=1.95, num_steps=S,
                hierarchical_sample=hierarchical, clamp_mode=clamp_mode, nerf_noise=nerf_noise,
                white_back=white_back, last_back=last_back,
                # extra curriculum keys the generator must ignore (configs/thousand/default.py)
                batch_size=B, gen_lr=5e-5, fade_steps=10000, z_lambda=0)
    film = oracle.SIREN_SPECS[oracle.resolve_siren_type(siren_type)].get("film", True)
    taps = run_reference_forward(siren_type, state, (vol, glob) if film else vol, cam, draws, meta)
    # parameters are NOT stored (MBs): they are regenerated from the seed by
    # oracle.init_generator_state; a float64 checksum detects a drifting torch CPU generator.
    fx = {"state/checksum": np.array(sum(float(v.double().abs().sum()) for v in state.values()))}
    fx.update({f"draw/{k}": v.numpy() for k, v in draws.items()})
    fx.update({f"tap/{k}": v.numpy() for k, v in taps.items()})
    fx["in/volume"], fx["in/global"], fx["in/cam2world"] = vol.numpy(), glob.numpy(), cam.numpy()
    extra = dict(siren_type=siren_type, seed=seed)
    if gains is not None:
        extra["dense_head_gains"] = gains
        w = taps["weights_final"][..., 0]
        print(f"  dense {siren_type}: far-plane weight mean {float(w[..., -1].mean()):.4f}, pixel std {float(taps['pixels'].std()):.3f}")
    fx["meta/json"] = np.array(__import__("json").dumps(dict(meta, **extra)))
    return fx


def make_function_fixture(seed=7):
    g = torch.Generator().manual_seed(seed)
    fx = {}
    # fancy_integration: [B,R,S,4]
    B, R, S = 2, 37, 20
    rs = torch.randn((B, R, S, 4), generator=g)
    rs[..., :3] = torch.sigmoid(rs[..., :3])
    rs[..., 3] *= 8
    t, _ = torch.sort(torch.rand((B, R, S, 1), generator=g) * 1.7 + 0.25, dim=-2)
    noise = torch.randn((B, R, S, 1), generator=g)
    fx["comp/rgb_sigma"], fx["comp/t"], fx["comp/noise"] = rs.numpy(), t.numpy(), noise.numpy()
    cases = [("relu", 0.0, False, False), ("relu", 0.7, True, False), ("softplus", 0.0, True, False),
             ("softplus", 0.3, False, True), ("relu", 0.0, True, True)]
    for i, (cm, ns, wb, lb) in enumerate(cases):
        rp = Replay([("randn", noise)])
        orig = torch.randn
        torch.randn = rp.randn
        try:
            rgb, dist, w = ref_vr.fancy_integration(rs, t, torch.device("cpu"), noise_std=ns, last_back=lb,
                                                    white_back=wb, clamp_mode=cm)
        finally:
            torch.randn = orig
        fx[f"comp/case{i}/cfg"] = np.array(__import__("json").dumps(dict(clamp_mode=cm, noise_std=ns, white_back=wb, last_back=lb)))
        fx[f"comp/case{i}/rgb"], fx[f"comp/case{i}/dist"], fx[f"comp/case{i}/weights"] = rgb.numpy(), dist.numpy(), w.numpy()
    # sample_pdf: [N, M+1] bins, [N, M] weights, K draws; rows 0..3 are degenerate
    N, M, K = 301, 22, 24
    bins, _ = torch.sort(torch.rand((N, M + 1), generator=g) * 1.7 + 0.25, dim=-1)
    w = torch.rand((N, M), generator=g) ** 4
    w[0] = 0.0                      # all-zero weights -> uniform pdf
    w[1] = 0.0; w[1, 5] = 1.0       # a single spike (many denom < eps bins)
    w[2, :11] = 0.0                 # leading zeros
    w[3] = 1e-5                     # tiny constant
    u = torch.rand((N, K), generator=g)
    u[4, 0], u[4, 1] = 0.0, 0.99999994   # extremes of U[0,1)
    rp = Replay([("rand", u)])
    orig = torch.rand
    torch.rand = rp.rand
    try:
        samples = ref_vr.sample_pdf(bins, w, K, det=False)
    finally:
        torch.rand = orig
    fx["pdf/bins"], fx["pdf/weights"], fx["pdf/u"], fx["pdf/samples"] = bins.numpy(), w.numpy(), u.numpy(), samples.numpy()
    # the reference does not return its indices; recompute them with the reference's own ops
    ww = w + 1e-5
    cdf = torch.cumsum(ww / torch.sum(ww, -1, keepdim=True), -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    fx["pdf/inds"] = torch.searchsorted(cdf, u.contiguous()).numpy()
    # F.grid_sample taps on scattered points incl. far outside the cube (border clamp)
    vol = torch.randn((1, 32, 9, 10, 11), generator=g)
    pts = (torch.rand((1, 16 * 16 * 4, 3), generator=g) - 0.5) * 2.0
    grid = (pts / 0.6).reshape(1, 16, 16, 4, 3)
    f = torch.nn.functional.grid_sample(vol, grid, mode="bilinear", align_corners=False, padding_mode="border")
    fx["tri/volume"], fx["tri/points"] = vol.numpy(), pts.numpy()
    fx["tri/features"] = f.reshape(1, 32, -1).permute(0, 2, 1).contiguous().numpy()
    return fx


class ReplayedRefGenerator(torch.nn.Module):
    """The unmodified reference generator with torch.rand / torch.randn replayed from ``metadata["draws"]`` on every forward."""

    def __init__(self, gen):
        super().__init__()
        self.gen = gen

    def forward(self, z, cam2worlds, **md):
        draws = md["draws"]
        meta = {k: v for k, v in md.items() if k != "draws"}
        rp = Replay([("rand", draws["u_jitter"]), ("randn", draws["noise_coarse"]), ("rand", draws["u_resample"]), ("randn", draws["noise_final"])])
        orig = torch.rand, torch.randn
        torch.rand, torch.randn = rp.rand, rp.randn
        try:
            return self.gen(z, cam2worlds, **meta)
        finally:
            torch.rand, torch.randn = orig


def make_train_step_fixture(steps=2):
    """Two optimisation steps of the reference's train step (oracle/train_step.py follows utils.py:621-842) on the REFERENCE's
    generator, U-Net and discriminator; also pins the U-Net / discriminator outputs themselves."""
    import types as _types
    tix = _types.ModuleType("tkinter.tix")
    tix.Tree = object
    sys.modules.setdefault("tkinter.tix", tix)
    import generators.unet3d as ref_unet
    import discriminators.discriminators as ref_disc
    from oracle import train_step as ts

    md = ts.tiny_config()
    gen = ref_gen.ImplicitGenerator3d(oracle.resolve_siren_type(ts.TINY_SIREN), z_dim=ts.TINY_ZDIM, input_dim=32, output_dim=4, hidden_dim=256)
    gen.load_state_dict(oracle.init_generator_state(ts.TINY_SIREN, z_dim=ts.TINY_ZDIM, seed=0), strict=True)
    gen.set_device(torch.device("cpu"))
    enc = ref_unet.UNet3D(**ts.TINY_UNET)
    disc = ref_disc.ProgressiveDiscriminator()
    ts.fill_params(enc, 1)
    ts.fill_params(disc, 2)
    sample = ts.tiny_sample()
    fx = {}
    with torch.no_grad():
        fv, glob = enc(sample["voxel"])
        fx["unet/volume"], fx["unet/global"] = fv.numpy(), glob.numpy()
        for size in (16, 64):
            img = torch.rand((2, 3, size, size), generator=torch.Generator().manual_seed(size)) * 2 - 1
            fx[f"disc/in{size}"], fx[f"disc/out{size}"] = img.numpy(), disc(img, 0.3).numpy()
        for cls, tag in ((ref_disc.ProgressiveEncoderDiscriminator, "enc"), (ref_disc.ProgressiveDiscriminator_inputCat, "cat")):
            d2 = cls()
            ts.fill_params(d2, 5)
            img = torch.tensor(fx["disc/in16"])
            out = d2(img, 0.3, cond=img.flip(0)) if tag == "cat" else torch.cat(d2(img, 0.3), dim=1)
            fx[f"disc_{tag}/out16"] = out.numpy()
        # the other two encoders of generators/unet3d.py (:829-898) on the same voxels
        pyr = ref_unet.PyramidUNet3D(**dict(ts.TINY_UNET, num_levels=3))
        res = ref_unet.ResidualUNet3D(**dict(ts.TINY_UNET, num_levels=3, return_global=False, out_channels=16))
        ts.fill_params(pyr, 3)
        ts.fill_params(res, 4)
        levels, pglob = pyr(sample["voxel"])
        for i, lv in enumerate(levels):
            fx[f"unet_pyramid/level{i}"] = lv.numpy()
        fx["unet_pyramid/global"] = pglob.numpy()
        fx["unet_residual/volume"] = res(sample["voxel"]).numpy()
    # curriculum helpers (configs/curriculums.py:83-137) on a four-stage curriculum shaped like configs/thousand/default.py
    from configs import curriculums as ref_cur
    cur = ts.tiny_curriculum()
    table = {str(st): [ref_cur.extract_metadata(cur, st)["img_size"], ref_cur.extract_metadata(cur, st)["batch_size"],
                       ref_cur.last_upsample_step(cur, st), float(min(ref_cur.next_upsample_step(cur, st), 1e9))]
             for st in (0, 1, 4999, 5000, 7000, 15000, 24999, 25000, 90000)}
    fx["curriculum/json"] = np.array(__import__("json").dumps(table))
    harness = ts.RefTrainStep(ReplayedRefGenerator(gen), enc, disc, dict(md, draws=ts.tiny_draws()), alpha=0.3)
    for i in range(steps):
        rec = harness.step(sample)
        for k, v in rec.items():
            fx[f"step{i}/{k}"] = np.array(v, dtype=np.float64)
        print(f"train step {i}:", {k: round(v, 6) for k, v in rec.items()})
    return fx


def make_latent_fixture(seed=41, img_size=16, S=12, B=2, z_dim=512):
    """``SHORTSIREN`` (siren.py:1172-1224; the default generator of configs/thousand/special.py:45-51): position input, latent z."""
    g = torch.Generator().manual_seed(seed)
    state = oracle.init_generator_state("SHORTSIREN", z_dim=z_dim, input_dim=3, hidden_dim=256, seed=seed)
    gains = [300.0, 3.0, 1.0, 6.0]                  # sigma gain, rgb gain, first-layer gain, sigma bias (thin fog: see oracle.DENSE_HEAD_GAINS)
    state = oracle.dense_head_state(state, *gains)
    latent = torch.randn((B, z_dim), generator=g)
    cam = oracle.look_at_cam2world(oracle.random_camera_origins(B, 0.7, 1.5, "y", np.random.RandomState(seed)), "y")
    draws = oracle.draw_randoms(B, img_size, S, True, g)
    meta = dict(img_size=img_size, fov=49.134342641202636, ray_start=0.25, ray_end=1.95, num_steps=S, hierarchical_sample=True,
                clamp_mode="relu", nerf_noise=0.0, white_back=True, last_back=False)
    gen = ref_gen.ImplicitGenerator3d("SHORTSIREN", z_dim=z_dim, input_dim=3, output_dim=4, hidden_dim=256)
    gen.load_state_dict(state, strict=True)
    gen.set_device(torch.device("cpu"))
    gen.eval()
    rp = Replay([("rand", draws["u_jitter"]), ("randn", draws["noise_coarse"]), ("rand", draws["u_resample"]), ("randn", draws["noise_final"])])
    calls = []
    orig_fwd = gen.siren.forward

    def tap(points, zz, *a):
        out = orig_fwd(points, zz, *a)
        calls.append((points.detach().clone(), out.detach().clone()))
        return out

    gen.siren.forward = tap
    orig = torch.rand, torch.randn
    torch.rand, torch.randn = rp.rand, rp.randn
    try:
        with torch.no_grad():
            pixels, depth = gen(latent, cam, **meta)
    finally:
        torch.rand, torch.randn = orig
    R = img_size ** 2
    fx = {"state/checksum": np.array(sum(float(v.double().abs().sum()) for v in state.values()))}
    fx.update({f"draw/{k}": v.numpy() for k, v in draws.items()})
    fx["in/latent"], fx["in/cam2world"] = latent.numpy(), cam.numpy()
    fx["tap/pixels"], fx["tap/depth"] = pixels.numpy(), depth.numpy()
    fx["tap/points_coarse"] = calls[0][0].reshape(B, R, S, 3).numpy()
    fx["tap/rgb_sigma_coarse"] = calls[0][1].reshape(B, R, S, 4).numpy()
    fx["tap/rgb_sigma_fine"] = calls[1][1].reshape(B, R, S, 4).numpy()
    fx["meta/json"] = np.array(__import__("json").dumps(dict(meta, siren_type="SHORTSIREN", seed=seed, z_dim=z_dim, dense_head_gains=gains)))
    return fx


LIBRARY_CASES = oracle.LIBRARY_CASES
library_inputs = oracle.library_case_inputs


def make_library_fixture(siren_type, img_size=12, S=8):
    z_dim, input_dim, plan, seed = LIBRARY_CASES[siren_type]
    state, z, cam, g = library_inputs(siren_type)
    B = cam.shape[0]
    draws = oracle.draw_randoms(B, img_size, S, True, g)
    meta = dict(img_size=img_size, fov=49.134342641202636, ray_start=0.25, ray_end=1.95, num_steps=S, hierarchical_sample=True,
                clamp_mode="softplus", nerf_noise=0.3, white_back=True, last_back=False)
    gen = ref_gen.ImplicitGenerator3d(siren_type, z_dim=z_dim, input_dim=input_dim, output_dim=4, hidden_dim=256)
    gen.load_state_dict(state, strict=True)
    gen.set_device(torch.device("cpu"))
    gen.eval()
    rp = Replay([("rand", draws["u_jitter"]), ("randn", draws["noise_coarse"]), ("rand", draws["u_resample"]), ("randn", draws["noise_final"])])
    calls = []
    orig_fwd = gen.siren.forward

    def tap(points, zz, *a):
        out = orig_fwd(points, zz, *a)
        calls.append((points.detach().clone(), out.detach().clone()))
        return out

    gen.siren.forward = tap
    orig = torch.rand, torch.randn
    torch.rand, torch.randn = rp.rand, rp.randn
    try:
        with torch.no_grad():
            pixels, depth = gen(z, cam, **meta)
    finally:
        torch.rand, torch.randn = orig
    R = img_size ** 2
    fx = {"state/checksum": np.array(sum(float(v.double().abs().sum()) for v in state.values()))}
    fx.update({f"draw/{k}": v.numpy() for k, v in draws.items()})
    fx["tap/pixels"], fx["tap/depth"] = pixels.numpy(), depth.numpy()
    fx["tap/points_coarse"] = calls[0][0].reshape(B, R, S, 3).numpy()
    fx["tap/rgb_sigma_coarse"] = calls[0][1].reshape(B, R, S, 4).numpy()
    fx["tap/rgb_sigma_fine"] = calls[1][1].reshape(B, R, S, 4).numpy()
    fx["meta/json"] = np.array(__import__("json").dumps(dict(meta, siren_type=siren_type)))
    return fx


def dense_fixtures():
    """Forward fixtures with real density: 16x16, 12+12 samples, 16^3 volume, batch 2."""
    kw = dict(img_size=16, S=12, V=16, B=2, dense=True)
    return {
        "fwd_dense_TALLSIREN_FG": make_forward_fixture("TALLSIREN_FG", 31, **kw),
        "fwd_dense_SHORTSIREN_FG": make_forward_fixture("SHORTSIREN_dg", 32, nerf_noise=0.5, **kw),
        "fwd_dense_DOUBLESIREN_FG": make_forward_fixture("DoubleSIREN_dg", 33, white_back=False, last_back=True, **kw),
        "fwd_dense_SingleSIREN_dg": make_forward_fixture("SingleSIREN_dg", 34, hierarchical=False, **kw),
    }


def main():
    out = {}
    if "--library-only" in sys.argv:
        for st in LIBRARY_CASES:
            np.savez_compressed(os.path.join(HERE, f"fwd_{st}.npz"), **make_library_fixture(st))
        return
    if "--latent-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "fwd_SHORTSIREN.npz"), **make_latent_fixture())
        return
    if "--dense-only" in sys.argv:
        for name, fx in dense_fixtures().items():
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
            print(f"{name}: {os.path.getsize(os.path.join(HERE, name + '.npz')) / 1024:.0f} KiB")
        return
    if "--dres-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "fwd_TALLSIREN_dRes.npz"), **make_forward_fixture("TALLSIREN_dRes", 16))
        np.savez_compressed(os.path.join(HERE, "fwd_TALLSIREN_dResLong.npz"), **make_forward_fixture("TALLSIREN_dResLong", 17, hierarchical=False))
        np.savez_compressed(os.path.join(HERE, "fwd_SHORTSIREN_FRes.npz"), **make_forward_fixture("SHORTSIREN_FRes", 18, clamp_mode="softplus", nerf_noise=0.3))
        return
    if "--train-only" in sys.argv:
        fx = make_train_step_fixture()
        np.savez_compressed(os.path.join(HERE, "train_step.npz"), **fx)
        return
    out["train_step"] = make_train_step_fixture()
    out["fwd_TALLSIREN_FG"] = make_forward_fixture("TALLSIREN_FG", 11)
    out["fwd_SHORTSIREN_FG"] = make_forward_fixture("SHORTSIREN_dg", 12, clamp_mode="softplus", nerf_noise=0.5, white_back=False, last_back=True)
    out["fwd_DOUBLESIREN_FG"] = make_forward_fixture("DoubleSIREN_dg", 13, hierarchical=False, white_back=True)
    out["fwd_SingleSIREN_dg"] = make_forward_fixture("SingleSIREN_dg", 14, nerf_noise=1.0)
    out["fwd_SHORTSIREN_F"] = make_forward_fixture("SHORTSIREN_F", 15)
    out["fwd_TALLSIREN_dRes"] = make_forward_fixture("TALLSIREN_dRes", 16)
    out["fwd_TALLSIREN_dResLong"] = make_forward_fixture("TALLSIREN_dResLong", 17, hierarchical=False)
    out["fwd_SHORTSIREN_FRes"] = make_forward_fixture("SHORTSIREN_FRes", 18, clamp_mode="softplus", nerf_noise=0.3)
    out["functions"] = make_function_fixture()
    out.update(dense_fixtures())
    out["fwd_SHORTSIREN"] = make_latent_fixture()
    for st in LIBRARY_CASES:
        out[f"fwd_{st}"] = make_library_fixture(st)
    for name, fx in out.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **fx)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(fx)} arrays")


if __name__ == "__main__":
    main()
