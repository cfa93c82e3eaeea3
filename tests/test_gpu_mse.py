"""GPU parity of decode and of the quantisation (MSE) loss + gradients against the golden fixtures
(reference autograd) and the oracle.  Floating point: tolerance 1e-4 relative (north_star)."""
import numpy as np
import pytest
import torch

from tests import golden_cases as GC
from tests.conftest import golden

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", list(GC.MSE_CASES))
def test_decode_exact_and_autograd(name):
    from repconc_b200 import decode
    case = GC.MSE_CASES[name]
    g = golden(name)
    x, c, gr, codes = GC.mse_inputs(case)
    ct = _dev(c).requires_grad_(True)
    q = decode(_dev(codes), ct)
    assert np.array_equal(q.detach().cpu().numpy(), g["quantized"])       # a gather is exact
    # the reference's own loss expression, autograd through OUR decode (scatter-add backward)
    xt = _dev(x).requires_grad_(True)
    gt = _dev(gr)
    surrogate = torch.dot(gt.flatten(), xt.flatten()) + torch.dot(gt.flatten(), q.flatten())
    mse = ((q - xt) ** 2).sum(-1).mean() * case["w"]
    (case["scale"] * mse + surrogate).backward()
    np.testing.assert_allclose(ct.grad.cpu().numpy(), g["grad_c"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(xt.grad.cpu().numpy(), g["grad_x"], rtol=RTOL, atol=1e-7)
    # numpy branch of decode, uint8 codes, transposed (non-contiguous) codes
    assert np.array_equal(decode(codes, ct), g["quantized"])
    assert np.array_equal(decode(_dev(codes.astype(np.uint8)), ct).detach().cpu().numpy(), g["quantized"])
    tview = _dev(np.ascontiguousarray(codes.T)).t()
    assert np.array_equal(decode(tview, ct).detach().cpu().numpy(), g["quantized"])


@pytest.mark.parametrize("name", list(GC.MSE_CASES))
def test_fused_quantization_loss(name, oracle):
    from repconc_b200.ops import quantization_loss
    case = GC.MSE_CASES[name]
    g = golden(name)
    x, c, gr, codes = GC.mse_inputs(case)
    xt = _dev(x).requires_grad_(True)
    ct = _dev(c).requires_grad_(True)
    mse, sur = quantization_loss(xt, ct, _dev(codes), _dev(gr), case["w"])
    np.testing.assert_allclose(mse.item(), g["mse"], rtol=RTOL)
    np.testing.assert_allclose(sur.item(), g["surrogate"], rtol=RTOL, atol=1e-6)
    (case["scale"] * mse + sur).backward()                  # finetune_repconc.py:390 / :396
    np.testing.assert_allclose(xt.grad.cpu().numpy(), g["grad_x"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(ct.grad.cpu().numpy(), g["grad_c"], rtol=RTOL, atol=1e-7)
    # oracle agreement on the same inputs
    o = oracle.mse_surrogate(x, oracle.decode(codes, c), gr, codes, case["K"], case["w"], case["scale"])
    np.testing.assert_allclose(xt.grad.cpu().numpy(), o["grad_x"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(ct.grad.cpu().numpy(), o["grad_c"], rtol=RTOL, atol=1e-7)


def test_loss_full_batch_against_oracle(oracle):
    """BASELINE config 3 shape: 8192 x 768, M=48 -- multi-chunk scatter-add, deterministic."""
    from repconc_b200.ops import quantization_loss
    r = np.random.default_rng(7)
    n, M, K, ds = 8192, 48, 256, 16
    x = r.standard_normal((n, M * ds), dtype=np.float32)
    c = r.standard_normal((M, K, ds), dtype=np.float32)
    gr = (r.standard_normal((n, M * ds), dtype=np.float32) / n).astype(np.float32)
    codes = r.integers(0, K, size=(n, M), dtype=np.int64)
    outs = []
    for _ in range(2):
        xt = _dev(x).requires_grad_(True)
        ct = _dev(c).requires_grad_(True)
        mse, sur = quantization_loss(xt, ct, _dev(codes), _dev(gr), 1e-4)
        (mse + sur).backward()
        outs.append((mse.item(), sur.item(), xt.grad.clone(), ct.grad.clone()))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][3], outs[1][3])      # deterministic
    o = oracle.mse_surrogate(x, oracle.decode(codes, c), gr, codes, K, 1e-4, 1.0)
    np.testing.assert_allclose(outs[0][0], o["mse"], rtol=RTOL)
    np.testing.assert_allclose(outs[0][1], o["surrogate"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(outs[0][2].cpu().numpy(), o["grad_x"], rtol=RTOL, atol=1e-8)
    np.testing.assert_allclose(outs[0][3].cpu().numpy(), o["grad_c"], rtol=RTOL, atol=1e-7)


def test_module_surface_on_gpu(oracle):
    """RepCONC module: quantize (both modes), decode, forward with a stub encoder."""
    from transformers import PretrainedConfig
    from repconc_b200 import RepCONC
    case = GC.ASSIGN_CASES["ds16_b512"]
    g = golden("assign_ds16_b512")
    x, c = GC.assign_inputs(case)
    cfg = PretrainedConfig(hidden_size=case["D"])
    cfg.MCQ_M, cfg.MCQ_K, cfg.similarity_metric = case["M"], case["K"], "METRIC_IP"

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.config = cfg
            self.table = torch.nn.Parameter(_dev(x).clone())

        def forward(self, input_ids, attention_mask):
            return self.table[input_ids[:, 0]]

    model = RepCONC(cfg, Enc(), True, case["eps"], case["iters"]).cuda()
    assert set(k.split(".")[0] for k in model.state_dict()) == {"rotation", "centroids", "dense_encoder"}
    with torch.no_grad():
        model.centroids.copy_(_dev(c))
    ids = torch.arange(case["B"], device="cuda")[:, None]
    out = model(ids, torch.ones_like(ids), return_code=True, return_quantized_embedding=True)
    assert np.array_equal(out.discrete_codes.cpu().numpy(), g["codes_conc"].astype(np.int64))
    assert np.array_equal(out.quantized_embeds.detach().cpu().numpy(),
                          oracle.decode(g["codes_conc"].astype(np.int64), c))
    model.use_constraint = False                       # callers flip this at run time
    assert np.array_equal(model.quantize(_dev(x)).cpu().numpy(), g["codes_nn"].astype(np.int64))
    out.quantized_embeds.sum().backward()
    assert model.centroids.grad is not None and model.centroids.grad.abs().sum() > 0


def test_decode_raises_on_out_of_range_code():
    """the reference's `centroids[first_indices, second_indices]` (modeling_repconc.py:171-175) raises IndexError on a
    code outside [0, K); rc_decode clamps and raises RC_FLAG_BADCODE, which ops.decode turns into the same error"""
    from repconc_b200 import ops
    c = torch.randn((8, 256, 16), device="cuda")
    codes = torch.randint(0, 256, (10, 8), device="cuda")
    ops.decode(codes, c)                                           # fine
    codes[3, 5] = 300
    with pytest.raises(IndexError):
        ops.decode(codes, c)
