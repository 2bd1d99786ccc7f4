"""GPU parity of K5 (fake-quant), K7 (bias-correction / cosine reductions) and the K6
elementwise kernels (soft rounding, fused gradient + Adam, epilogues) through the C-ABI,
against the oracle (ONNX Q/DQ semantics, torch autograd + torch.optim.Adam in fp32)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fakequant_known_answers(dpl_built):
    """round-half-even, int8 saturation at [-128, 127] for ONNX Q/DQ and
    [-127, 127] for quant_acti."""
    import torch
    from dipoorlet_b200 import kernels as K
    s = torch.tensor([0.5], device="cuda")
    x = torch.tensor([0.25, 0.75, 1.25, -0.25, -63.8, 100.0, -100.0, 0.0], device="cuda")  # x/s = .5 1.5 2.5 -.5 -127.6 200 -200
    y = K.fakequant(x, s, None, -128, 127).cpu().numpy()
    assert np.array_equal(y, np.array([0.0, 1.0, 1.0, -0.0, -64.0, 63.5, -64.0, 0.0], np.float32))
    y = K.fakequant(x, s, None, -127, 127).cpu().numpy()
    assert y[4] == -63.5 and y[6] == -63.5
    zp = torch.tensor([128], dtype=torch.int32, device="cuda")
    y = K.fakequant(x, s, zp, 0, 255).cpu().numpy()   # uint8, zero point 128
    assert np.array_equal(y, np.array([0.0, 1.0, 1.0, -0.0, -64.0, 63.5, -64.0, 0.0], np.float32))


@pytest.mark.parametrize("per_channel", [False, True])
def test_fakequant_matches_onnx_qdq(dpl_built, per_channel):
    import torch
    from dipoorlet_b200 import kernels as K
    from dipoorlet_b200 import onnx_lite as ol
    from oracle import forward as OF
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((6, 8, 5, 7)) * 3).astype(np.float32)
    scale = (rng.random(8 if per_channel else 1) * 0.05 + 0.01).astype(np.float32)
    zp = np.zeros_like(scale, dtype=np.int8)
    attrs = {"axis": 1} if per_channel else {}
    qn = ol.Node("QuantizeLinear", ["x", "s", "z"], ["q"], "q", attrs)
    dn = ol.Node("DequantizeLinear", ["q", "s", "z"], ["y"], "dq", attrs)
    t = lambda a: torch.from_numpy(a)  # noqa: E731
    q = OF.run_node(qn, [t(x), t(scale if per_channel else scale.reshape(())), t(zp if per_channel else zp.reshape(()))], attrs)
    want = OF.run_node(dn, [q, t(scale if per_channel else scale.reshape(())), t(zp if per_channel else zp.reshape(()))], attrs).numpy()
    got = K.fakequant(t(x).cuda(), t(scale).cuda(), None, -128, 127, axis=1 if per_channel else None).cpu().numpy()
    assert np.array_equal(got, want)


def test_fakequant_rounding_ties_bit_exact(dpl_built):
    """The quotient is formed as x * RN(1/s) + one FMA correction, with the exact IEEE division only
    next to a rounding tie: adversarial inputs on and one ulp around every tie, for awkward scales,
    must round exactly like round_half_even(x / s) in IEEE fp32 (NumPy), and so must random data."""
    import torch
    from dipoorlet_b200 import kernels as K
    rng = np.random.default_rng(5)
    scales = np.concatenate([rng.random(40).astype(np.float32) * 0.2 + 1e-3,
                             np.array([1 / 3, 0.1, 0.7, 1e-6, 3e4, np.float32(2 ** -10), 0.99999994], np.float32)])
    n = np.arange(-140, 141, dtype=np.float32)
    for s in scales:
        ties = ((n + np.float32(0.5)) * s).astype(np.float32)
        xs = [ties]
        for k in range(1, 4):
            up, dn = ties, ties
            for _ in range(k):
                up, dn = np.nextafter(up, np.float32(np.inf)), np.nextafter(dn, np.float32(-np.inf))
            xs += [up.astype(np.float32), dn.astype(np.float32)]
        x = np.concatenate(xs + [(rng.standard_normal(4096) * 60 * s).astype(np.float32),
                                 np.array([np.inf, -np.inf, 0.0, -0.0, 3e38, -3e38], np.float32)])
        q = np.rint(x / np.float32(s))                       # IEEE fp32 division, half-even
        for lo, hi in ((-128, 127), (-127, 127)):
            want = (np.clip(q, lo, hi) * np.float32(s)).astype(np.float32)
            got = K.fakequant(torch.from_numpy(x).cuda(), torch.tensor([s], device="cuda"), None, lo, hi).cpu().numpy()
            assert np.array_equal(got, want), (s, lo, np.flatnonzero(got != want)[:5])
        # the reconstruction loop's epilogue shares the quotient (quant_acti, ada_quant_layer.py:28-36)
        got = K.recon_act(torch.from_numpy(x).cuda(), False, (float(s), -127.0, 127.0)).cpu().numpy()
        assert np.array_equal(got, (np.clip(q, -127, 127) * np.float32(s)).astype(np.float32)), s


def test_qdrop_mask_is_a_function_of_seed_and_index(dpl_built):
    """The Bernoulli mask (brecq.py:169-170, ada_quant_layer.py:28-36) is regenerated, never stored:
    the 16-byte and the scalar code paths, and the forward / backward kernels, must agree on it, its
    mean must be the drop probability and different seeds must give different masks."""
    import torch
    from dipoorlet_b200 import kernels as K
    n = 1 << 20
    ones, zeros = torch.ones(n + 4, device="cuda"), torch.zeros(n + 4, device="cuda")
    for p in (0.5, 0.3):
        m_vec = K.mix_drop(ones[:n], zeros[:n], p, 11)
        # misaligned operands take the scalar path over the same element indices
        m_sca = K.mix_drop(ones[1:n + 1], zeros[1:n + 1], p, 11, out=torch.empty(n + 1, device="cuda")[1:])
        assert torch.equal(m_vec, m_sca)
        assert abs(m_vec.mean().item() - p) < 4 * (p * (1 - p) / n) ** 0.5
        assert not torch.equal(m_vec, K.mix_drop(ones[:n], zeros[:n], p, 12))
        # lag-1 correlation of neighbours (two elements share one hash word)
        a, b = m_vec[:-1] - p, m_vec[1:] - p
        assert abs((a * b).mean().item()) < 5e-3
    o = torch.randn(n, device="cuda") * 3
    gy = torch.ones(n, device="cuda")
    y = K.recon_act(o, False, (0.05, -127.0, 127.0), prob=0.5, seed=5)
    go = K.recon_act_bwd(o, gy, False, (0.05, -127.0, 127.0), prob=0.5, seed=5)
    quantised = y != o                      # (a quantised value can equal o only on exact grid points)
    assert torch.equal(go == 0, quantised | ((go == 0) & (y == o)))
    assert (go[quantised] == 0).all() and abs(quantised.float().mean().item() - 0.5) < 0.01
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    gl = K.recon_loss(o, torch.zeros_like(o), 1.0, loss, False, (0.05, -127.0, 127.0), prob=0.5, seed=5)
    assert torch.equal(gl == 0, go == 0) or ((gl == 0) != (go == 0)).float().mean().item() < 1e-3


def test_channel_sumdiff_and_cosine(dpl_built):
    import torch
    from dipoorlet_b200 import kernels as K
    rng = np.random.default_rng(1)
    for shape in [(5, 7, 9, 4), (6, 10)]:
        a = rng.standard_normal(shape).astype(np.float32)
        b = (a + 0.01 * rng.standard_normal(shape)).astype(np.float32)
        acc = torch.zeros(shape[1], dtype=torch.float64, device="cuda")
        K.channel_sumdiff(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), shape[1], acc)
        axis = (0, 2, 3) if len(shape) == 4 else 0
        want = (a.astype(np.float64) - b).sum(axis=axis)
        assert np.allclose(acc.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
        out = torch.zeros((shape[0], 3), dtype=torch.float64, device="cuda")
        K.cosine3(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), out)
        a2, b2 = a.reshape(shape[0], -1).astype(np.float64), b.reshape(shape[0], -1).astype(np.float64)
        want = np.stack([(a2 * b2).sum(1), (a2 * a2).sum(1), (b2 * b2).sum(1)], 1)
        assert np.allclose(out.cpu().numpy(), want, rtol=1e-6)


def test_adaround_init_and_weight(dpl_built):
    """h(alpha0) == frac(w/s) (SURVEY.md A-8 KAT); soft/hard weights == the torch expressions."""
    import torch
    from dipoorlet_b200 import kernels as K
    from oracle import adaround as OA
    g = torch.Generator(device="cuda").manual_seed(0)
    w = torch.randn((16, 8, 3, 3), device="cuda", generator=g) * 0.2
    scale = (w.abs().amax(dim=(1, 2, 3)) / 127).contiguous()
    s4 = scale.view(-1, 1, 1, 1)
    alpha, wfloor = K.adaround_init(w, scale)
    assert torch.equal(wfloor, (w / s4).floor())
    assert torch.allclose(alpha, OA.alpha_init(w, s4), rtol=1e-5, atol=1e-6)
    rest = (w / s4) - (w / s4).floor()
    assert torch.allclose(OA.rectified_sigmoid(alpha), rest, atol=2e-6)
    qmin, qmax = torch.full_like(s4, -127), torch.full_like(s4, 127)
    a2 = alpha + torch.randn(alpha.shape, device="cuda", generator=g)
    for soft in (True, False):
        got = K.adaround_weight(wfloor, a2, scale, -127, 127, soft)
        want = OA.quant_weight(w, a2, s4, qmin, qmax, soft)
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7), soft
    hard = K.adaround_weight(wfloor, a2, scale, -127, 127, False)
    assert torch.equal(hard, torch.clamp((w / s4).floor() + (a2 >= 0).float(), -127, 127) * s4)


@pytest.mark.parametrize("relu,drop", [(True, False), (False, False), (True, True)])
def test_learning_loop_matches_torch_autograd(dpl_built, relu, drop, monkeypatch):
    """A 2-layer block (conv3x3 -> [relu] -> conv1x1) for 30 iterations: the fused launch
    sequence must track torch autograd + torch.optim.Adam (same device, fp32)."""
    import torch
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200.weight_transform.ada_quant_layer import AdaQLayer, adaround_reg
    from dipoorlet_b200.weight_transform.learning import learning_round_mask
    from oracle import adaround as OA
    monkeypatch.setenv("DPL_RECON_TF32", "0")   # fp32 vs fp32: this test is about the update rule
    monkeypatch.setenv("DPL_CUDA_GRAPH_MIN_ITERS", "8")   # exercise the captured path
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    n, bs, epochs = 16, 8, 15
    x = torch.randn((n, 6, 10, 10), device=dev, generator=g)
    w1 = torch.randn((8, 6, 3, 3), device=dev, generator=g) * 0.2
    b1 = torch.randn(8, device=dev, generator=g) * 0.1
    w2 = torch.randn((5, 8, 1, 1), device=dev, generator=g) * 0.3
    attrs1 = {"dilations": [1, 1], "group": 1, "kernel_shape": [3, 3], "pads": [1, 1, 1, 1], "strides": [1, 1]}
    attrs2 = {"dilations": [1, 1], "group": 1, "kernel_shape": [1, 1], "pads": [0, 0, 0, 0], "strides": [2, 2]}
    # the block input of the quantised graph differs from the fp one (as in the real flow):
    # with identical inputs the initial loss is exactly zero (h(alpha0) = frac(w/s)) and the
    # first Adam steps would be driven by rounding noise below eps, i.e. implementation chaos
    x_fp = x
    x = torch.round(x_fp / 0.1) * 0.1
    with torch.no_grad():
        h = torch.nn.functional.conv2d(x_fp, w1, b1, padding=1)
        h = torch.relu(h) if relu else h
        tgt = torch.nn.functional.conv2d(h, w2, None, stride=2)
    total_iter = epochs * 2 * np.ceil(n / bs)
    # reg active from the start for a stronger test: total_iter small => beta > 0 after 20 %
    specs = [(attrs1, w1, b1, relu), (attrs2, w2, None, False)]
    ref_layers, layers = [], []
    for a, w, b, r in specs:
        scale = (w.abs().amax(dim=(1, 2, 3)) / 127).contiguous()
        s4 = scale.view(-1, 1, 1, 1)
        qi = (torch.tensor(0.05, device=dev), torch.tensor(-127., device=dev), torch.tensor(127., device=dev))
        ref_layers.append(OA.Layer("Conv", a, w, b, s4, torch.full_like(s4, -127), torch.full_like(s4, 127), r,
                                   qi=qi, acti_quant=False))
        node = ol.Node("Conv", ["x", "w"], ["y"], "c", a)
        layers.append(AdaQLayer(node, w, b, scale, -127, 127, r, qi=(0.05, -127., 127.), acti_quant=drop, device=dev))
    reg = adaround_reg(total_iter)
    if drop:
        # QDrop (brecq.py:167-172, ada_quant_layer.py:28-36,247-249): the kernels draw their Bernoulli masks
        # from a counter-based hash, torch from Philox, so the reference loop below is fed the kernels' own
        # masks: the per-epoch input mix is taken from dpl_mix_drop_f32 directly, the per-iteration output
        # masks are regenerated from the schedule kernel's seeds with a probe tensor. Everything else is the
        # reference's arithmetic under torch autograd + torch.optim.Adam.
        import torch.nn.functional as F
        from dipoorlet_b200 import kernels as K
        from dipoorlet_b200.weight_transform import learning
        seed = 5
        for l in ref_layers:
            l.acti_quant = True
        n_ep, n_b = epochs * 2, int(np.ceil(n / bs))
        d_iter = torch.zeros(1, dtype=torch.int32, device=dev)
        sched = torch.zeros(4, dtype=torch.float32, device=dev)
        seeds = torch.zeros(len(layers) + 1, dtype=torch.int64, device=dev)
        opt = torch.optim.Adam([l.round_mask for l in ref_layers])
        cur = 0
        for ep in range(n_ep):
            in_tensor = K.mix_drop(x, x_fp, 0.5, learning._seed(seed, ep, 991))
            for idx in range(n_b):
                K.recon_schedule(d_iter, sched, seeds, float(total_iter), seed_base=seed)
                out = in_tensor[idx * bs:(idx + 1) * bs]
                for li, l in enumerate(ref_layers):
                    wq = OA.quant_weight(l.weight, l.round_mask, l.scale, l.q_min, l.q_max)
                    a = l.attrs
                    out = F.conv2d(out, wq, l.bias, a["strides"], a["pads"][:2], a["dilations"], a["group"])
                    if l.relu_flag:
                        out = F.relu(out)
                    probe = torch.full_like(out, 0.3).detach()
                    quantised = K.recon_act(probe, relu=False, quant=(1.0, -127., 127.), prob=0.5,
                                            seed_dev=seeds[li:li + 1]) == 0
                    oq = torch.clamp((out / l.qi[0]).round(), -127, 127) * l.qi[0]
                    out = torch.where(quantised, oq, out)
                l2 = OA.l2_norm(out, tgt[idx * bs:(idx + 1) * bs])
                loss = l2
                beta = OA.temp_decay(cur, total_iter)
                for l in ref_layers:
                    loss = loss + OA.reg_loss(l.round_mask, beta)
                cur += 1
                opt.zero_grad()
                loss.backward()
                opt.step()
        monkeypatch.setenv("DPL_CUDA_GRAPH", "0")
        got_loss = learning_round_mask(layers, x, tgt, reg, bs, n_ep, fp_in=x_fp, drop=True, seed=seed)
        assert abs(got_loss - float(l2)) <= 2e-3 * max(1.0, abs(float(l2))), (got_loss, float(l2))
    else:
        OA.learn(ref_layers, x, tgt, total_iter, bs, epochs * 2)
        loss = learning_round_mask(layers, x, tgt, reg, bs, epochs * 2)
    for ref, got in zip(ref_layers, layers):
        # The largest weight of every channel has w/s = +-127 exactly, so h(alpha0) sits ON the
        # clamp boundary 0: whether the gradient passes there depends on the last bit of
        # sigmoid(), and Adam turns any non-zero gradient into a full +-lr step. Those
        # (n_channels) elements are implementation-defined in the reference too: exclude them.
        q = ref.weight / ref.scale
        interior = ((q - q.floor()) > 1e-4) & ((q - q.floor()) < 1 - 1e-4)
        d = (ref.round_mask.detach() - got.round_mask).abs()[interior].max().item()
        assert d < 2e-4, d
        assert interior.float().mean().item() > 0.8
        same = ((ref.round_mask.detach() >= 0) == (got.round_mask >= 0))[interior].float().mean().item()
        assert same > 0.999


def test_cuda_graph_replay_equals_eager(dpl_built, monkeypatch):
    """The captured-and-replayed iteration (DPL_CUDA_GRAPH=1) must produce the same alpha as the
    eager launch sequence, including a ragged last mini-batch that runs eagerly in both."""
    import torch
    from dipoorlet_b200 import onnx_lite as ol
    from dipoorlet_b200.weight_transform.ada_quant_layer import AdaQLayer, adaround_reg
    from dipoorlet_b200.weight_transform.learning import learning_round_mask
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(7)
    n, bs, epochs = 20, 8, 10          # batches of 8, 8, 4
    x_fp = torch.randn((n, 8, 12, 12), device=dev, generator=g)
    x = torch.round(x_fp / 0.05) * 0.05
    w = torch.randn((16, 8, 3, 3), device=dev, generator=g) * 0.2
    b = torch.randn(16, device=dev, generator=g) * 0.1
    attrs = {"dilations": [1, 1], "group": 1, "kernel_shape": [3, 3], "pads": [1, 1, 1, 1], "strides": [1, 1]}
    with torch.no_grad():
        tgt = torch.relu(torch.nn.functional.conv2d(x_fp, w, b, padding=1))
    scale = (w.abs().amax(dim=(1, 2, 3)) / 127).contiguous()
    res = []
    monkeypatch.setenv("DPL_CUDA_GRAPH_MIN_ITERS", "8")
    for mode in ("0", "1"):
        monkeypatch.setenv("DPL_CUDA_GRAPH", mode)
        layer = AdaQLayer(ol.Node("Conv", ["x", "w"], ["y"], "c", attrs), w, b, scale, -127, 127, True, device=dev)
        reg = adaround_reg(epochs * 3)
        learning_round_mask([layer], x, tgt, reg, bs, epochs, seed=3)
        res.append(layer.round_mask.clone())
    assert torch.allclose(res[0], res[1], rtol=0, atol=1e-6), (res[0] - res[1]).abs().max().item()
    assert (res[0] - AdaQLayer(ol.Node("Conv", ["x", "w"], ["y"], "c", attrs), w, b, scale, -127, 127, True,
                               device=dev).round_mask).abs().max().item() > 1e-3   # it did learn
