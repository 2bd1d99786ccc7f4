"""K6 step with the gradient all-reduce inside (dpl_adaround_step_peer_f32), on ONE GPU:
 * world = 1 must be dpl_adaround_step_f32 bit for bit;
 * world = 2 emulated by two "ranks" on two streams of the same device (peer pointers are plain device
   pointers there): both kernels must meet through the arrival words and produce bit-identical replicas equal
   to the plain step on the rank-ordered sum.
Verified on hardware in round 2 (one GPU: these tests; 2 and 8 GPUs: tools/multi_gpu_finetune.sh,
profiles/r2_finetune_multi_gpu.json - bit-identical replicas). DPL_PEER_ALLREDUCE=1 stays opt-in because it is
not faster than NCCL's all-reduce at these sizes."""
import os

import pytest

pytestmark = pytest.mark.gpu


def _state(n, c, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(n, device="cuda", generator=g)
    scale = torch.rand(c, device="cuda", generator=g) * 0.05 + 0.01
    alpha = torch.randn(n, device="cuda", generator=g)
    return w.floor(), scale, alpha, torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")


@pytest.mark.parametrize("n,c", [(64 * 576, 64), (1000 * 3, 1000), (7, 1)])
def test_world1_equals_plain_step(dpl_built, n, c):
    import torch
    from dipoorlet_b200 import kernels as K
    wfloor, scale, alpha, m, v = _state(n, c, 1)
    grad = torch.randn(n, device="cuda")
    a2, m2, v2 = alpha.clone(), m.clone(), v.clone()
    words = torch.zeros(64, dtype=torch.int32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    for step in (1, 2, 3):
        K.adaround_step(grad, wfloor, scale, -127, 127, 8.0, alpha, m, v, step)
        K.adaround_step_peer([grad.data_ptr()], [words.data_ptr()], 0, step, wfloor, scale, -127, 127, 8.0,
                             a2, m2, v2, step, error=err)
    torch.cuda.synchronize()
    assert int(err.item()) == 0 and int(words[0].item()) == 3
    assert torch.equal(alpha, a2) and torch.equal(m, m2) and torch.equal(v, v2)


def test_two_ranks_on_two_streams(dpl_built):
    import torch
    from dipoorlet_b200 import kernels as K
    n, c = 256 * 1152, 256
    wfloor, scale, alpha, m, v = _state(n, c, 2)
    grads = [torch.randn(2, n, device="cuda") for _ in range(2)]        # [rank][slot]
    words = [torch.zeros(64, dtype=torch.int32, device="cuda") for _ in range(2)]
    reps = [(alpha.clone(), m.clone(), v.clone()) for _ in range(2)]
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()
    for epoch in (1, 2, 3, 4):
        ptr_g = [g[epoch & 1].data_ptr() for g in grads]
        ptr_w = [w.data_ptr() for w in words]
        for rank in (1, 0):                      # launch order must not matter
            with torch.cuda.stream(streams[rank]):
                a, mm, vv = reps[rank]
                K.adaround_step_peer(ptr_g, ptr_w, rank, epoch, wfloor, scale, -127, 127, 8.0, a, mm, vv, epoch,
                                     error=err)
        # reference: the plain step on the rank-ordered sum, mean folded in
        K.adaround_step(grads[0][epoch & 1] + grads[1][epoch & 1], wfloor, scale, -127, 127, 8.0, alpha, m, v,
                        epoch, grad_scale=0.5)
        torch.cuda.synchronize()
    assert int(err.item()) == 0
    for a, mm, vv in reps:
        assert torch.equal(a, alpha) and torch.equal(mm, m) and torch.equal(vv, v)
