"""GPU parity: the whole inference graph (stem s2d, fused C3, SPPF pool, up-sample copies, concat-offset
stores, Detect decode) against (1) golden outputs of the REFERENCE model and (2) the fp32 torch oracle.

Tolerance: activations are bf16 between layers (the reference's --half path uses fp16), so the decoded
output is held to |err| <= 3e-2*|ref| + 3e-2 element-wise with a mean error below 4e-3; theta argmax must
agree on > 97 % of anchors (ties between neighbouring CSL bins flip under any rounding)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import model_ref
from tests.modelgen import build_mirror

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
DEV = "cuda:0"


def _compare(got, ref, what):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    tol = ref.abs() * 3e-2 + 3e-2
    frac_bad = (err > tol).float().mean().item()
    print(f"{what}: max err {err.max().item():.4g}, mean err {err.mean().item():.4g}, out-of-tol {frac_bad:.2e}")
    assert frac_bad < 1e-3, f"{what}: {frac_bad:.3e} of elements out of tolerance (max {err.max().item():.4g})"
    assert err.mean().item() < 4e-3
    am_g, am_r = got[..., 20:].argmax(-1), ref[..., 20:].argmax(-1)
    assert (am_g == am_r).float().mean().item() > 0.97


@pytest.mark.parametrize("size", ["n", "s"])
def test_engine_matches_reference_golden(size):
    G = np.load(ROOT / "tests" / "golden" / "model_golden.npz")
    m = build_mirror(size, nc=15, seed=0).to(DEV)
    x = torch.from_numpy(G[f"{size}/x"]).to(DEV)
    pred, _ = m(x)
    torch.cuda.synchronize()
    assert pred.shape == (1, 378, 200) and pred.dtype == torch.float32
    _compare(pred, torch.from_numpy(G[f"{size}/pred"]), f"yolov5{size} vs reference golden")


@pytest.mark.parametrize("size,B,H,W", [("n", 2, 128, 160), ("m", 1, 96, 64), ("x", 1, 64, 64)])
def test_engine_matches_fp32_oracle(size, B, H, W):
    m = build_mirror(size, nc=15, seed=1)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 3, H, W, generator=g)
    ref, _ = model_ref.forward(m, x)
    md = m.to(DEV)
    pred, _ = md(x.to(DEV))
    torch.cuda.synchronize()
    _compare(pred, ref, f"yolov5{size} {B}x{H}x{W} vs fp32 oracle")
    # a second call reuses the plan and is deterministic
    p1 = pred.clone()
    pred2, _ = md(x.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(p1, pred2)


def test_engine_refuses_what_it_does_not_run():
    m = build_mirror("n", nc=15, seed=0).to(DEV)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))          # CPU tensor
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 72, 64, device=DEV))  # not a stride multiple
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64, device=DEV), augment=True)  # TTA is host tooling outside the path
    m.train()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))          # training mode has no CPU path either


def test_bench_plan_yolov5s_b16_1024_matches_oracle():
    """The EXACT plan bench.py times (yolov5s, b16 x 1024^2: row-shift 8x16 tiles on the 256^2 / 128^2 maps, stride-2 pixel-pair
    views 512 wide, the stem's 514-pixel padded rows, 16 images per grid) against the fp32 oracle on three sampled images of
    the batch: relative L2 per Detect level < 3e-2 and the `obj > conf` candidate masks agree (bench.parity_gate, which
    bench.py also runs before it times anything)."""
    import copy
    import bench
    m_cpu = bench.build_model("s")
    m = copy.deepcopy(m_cpu).to(DEV)
    x = bench.synth_batch(16, 0).to(DEV)
    pred, _ = m(x)
    torch.cuda.synchronize()
    out = bench.parity_gate(m_cpu, x, pred, sample=(0, 5, -1))
    print(out)
    assert out["ok"]
    # graph replay (third call on) returns the same bits as the eager launches
    p1 = pred.clone()
    for _ in range(3):
        pred, _ = m(x)
    torch.cuda.synchronize()
    assert torch.equal(p1, pred)


def test_bench_plan_yolov5m_b16_1024_matches_oracle():
    """north_star's yolov5m b16 inference shape, one sampled image against the fp32 oracle."""
    import copy
    import bench
    m_cpu = bench.build_model("m")
    m = copy.deepcopy(m_cpu).to(DEV)
    x = bench.synth_batch(16, 0).to(DEV)
    pred, _ = m(x)
    torch.cuda.synchronize()
    out = bench.parity_gate(m_cpu, x, pred, sample=(3,))
    print(out)
    assert out["ok"]


@pytest.mark.parametrize("size,B", [("s", 16), ("m", 16)])
def test_bench_plan_teacher_forced(size, B):
    """Chaos-free, layer-by-layer parity of the benchmarked plan (b16 x 1024^2): every top-level module (Conv, C3 with its fused
    cv1|cv2 GEMM and in-place Bottleneck chain, SPPF) of the fp32 oracle is evaluated on the DEVICE's own input of that layer
    (two sampled images) and compared with what the device produced: relative L2 < 1.5e-2 (bf16 storage, 2-7 convolutions deep).
    A tile-geometry bug of any layer at this shape shows up here whatever the network's sensitivity."""
    import copy
    import bench
    import yolov5_obb_b200.yolo as Y
    m_cpu = bench.build_model(size)
    m = copy.deepcopy(m_cpu).to(DEV)
    x = bench.synth_batch(B, 0).to(DEV)
    m(x)
    torch.cuda.synchronize()
    eng = m._engines[(tuple(x.shape), 0)]
    imgs = [0, B - 1]

    def nchw(s):
        return s.buf[imgs][..., s.c_off:s.c_off + s.C].float().permute(0, 3, 1, 2).contiguous().cpu()

    worst = 0.0
    for i, mod in enumerate(m_cpu.model):
        if not isinstance(mod, (Y.Conv, Y.C3, Y.SPPF)):
            continue
        xin = (x[imgs].float() / 255).cpu().bfloat16().float() if i == 0 else nchw(eng.out_slices[i - 1 if mod.f == -1 else mod.f])
        fn = {Y.Conv: model_ref.conv_fwd, Y.C3: model_ref.c3_fwd, Y.SPPF: model_ref.sppf_fwd}[type(mod)]
        with torch.no_grad():
            want = fn(mod, xin, False)
        have = nchw(eng.out_slices[i])
        rel = ((have - want).norm() / want.norm()).item()
        worst = max(worst, rel)
        assert rel < 1.5e-2, (i, type(mod).__name__, rel)
    print(f"yolov5{size} b{B} 1024^2: worst teacher-forced module deviation {worst:.3e}")
    # Detect: decoded rows of the two images against the oracle's Detect on the device's own P3/P4/P5 inputs
    det = m_cpu.model[-1]
    with torch.no_grad():
        want, _ = model_ref.detect_fwd(det, [nchw(eng.out_slices[f]) for f in det.f], False)
    got = eng.pred[imgs].float().cpu()
    rel = ((got - want).norm() / want.norm()).item()
    prob_err = (got[..., 4:] - want[..., 4:]).abs().max().item()
    print(f"Detect on the device's own inputs: rel L2 {rel:.3e}, max |d prob| {prob_err:.3e}")
    assert rel < 1e-2 and prob_err < 5e-2
