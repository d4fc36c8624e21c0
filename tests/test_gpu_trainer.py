"""GPU tests of the optimizer tail and the graph-captured training step."""
import numpy as np
import pytest
import torch

from oracle import siren_oracle as so
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("clip", [0, 1])
def test_fused_adam_matches_torch_adam(clip, golden_dir):
    import os
    from siren_mri_b200.optim import FusedAdam
    g = np.load(os.path.join(golden_dir, "adam_clip%d.npz" % clip))
    p = torch.from_numpy(g["p0"].copy()).cuda()
    grad = torch.zeros_like(p)
    opt = FusedAdam(p, grad, lr=float(g["lr"]), max_grad_norm=float(g["clip"]))
    for i, gr in enumerate(g["grads"]):
        grad.copy_(torch.from_numpy(gr))
        opt.step()
        assert np.abs(p.cpu().numpy() - g["traj"][i]).max() < 2e-6, i


def _oracle_steps(Ws, bs, x, gt, steps, lr):
    W = [w.astype(np.float64) for w in Ws]
    b = [v.astype(np.float64) for v in bs]
    mW = [np.zeros_like(w) for w in W]; vW = [np.zeros_like(w) for w in W]
    mb = [np.zeros_like(v) for v in b]; vb = [np.zeros_like(v) for v in b]
    losses = []
    for s in range(1, steps + 1):
        y, _, _, cache = so.siren_forward(x.astype(np.float64), W, b, 30.0, 0)
        loss, gy = so.image_mse(y, gt.astype(np.float64))
        losses.append(loss)
        dW, db, _ = so.siren_backward(cache, W, gy)
        for l in range(len(W)):
            W[l], mW[l], vW[l] = so.adam_step(W[l], dW[l], mW[l], vW[l], s, lr=lr)
            b[l], mb[l], vb[l] = so.adam_step(b[l], db[l], mb[l], vb[l], s, lr=lr)
    return W, b, losses


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_matches_oracle_after_1_and_10_steps(graph):
    from siren_mri_b200 import modules
    from siren_mri_b200.trainer import SirenTrainer
    n = 3000
    Ws, bs = so.make_params(2, 256, 3, 1, seed=21)
    x = so.make_coords(1, n, 2, seed=22)
    gt = np.random.default_rng(23).uniform(-1, 1, size=(1, n, 1)).astype(np.float32)
    m = modules.SingleBVPNet(in_features=2, out_features=1, precision="fp32").cuda()
    with torch.no_grad():
        for l in range(5):
            m.net.net[l][0].weight.copy_(torch.from_numpy(Ws[l]))
            m.net.net[l][0].bias.copy_(torch.from_numpy(bs[l]))
    tr = SirenTrainer(m, n, lr=1e-4, use_graph=graph)
    tr.coords.copy_(torch.from_numpy(x))
    tr.gt.copy_(torch.from_numpy(gt))
    for steps in (1, 10):
        W_o, b_o, losses = _oracle_steps(Ws, bs, x, gt, steps, 1e-4)
        while tr.steps < steps:
            tr.step()
        torch.cuda.synchronize()
        for l in range(5):
            got = m.net.net[l][0].weight.detach().cpu().numpy()
            # compare the UPDATE (w - w0): Adam's first steps are +-lr per element
            du, do = got - Ws[l], W_o[l] - Ws[l]
            assert rel_l2(du, do) < 2e-2, (steps, l, rel_l2(du, do))
            assert rel_l2(got, W_o[l]) < 1e-4     # a handful of +-lr sign flips where |g| ~ 0
        assert abs(float(tr.loss.item()) - losses[-1]) < 1e-4 * abs(losses[-1])
    # the module still works through the ordinary forward after its parameters were re-homed
    with torch.no_grad():
        y = m({"coords": torch.from_numpy(x).cuda()})["model_out"]
    assert torch.isfinite(y).all()


def test_step_from_host_returns_loss():
    from siren_mri_b200 import modules
    from siren_mri_b200.trainer import SirenTrainer
    n = 1024
    m = modules.SingleBVPNet(in_features=2, out_features=1, precision="bf16").cuda()
    tr = SirenTrainer(m, n, lr=1e-4)
    xs = torch.rand(1, n, 2).pin_memory()
    gt = torch.rand(1, n, 1).pin_memory()
    l0 = tr.step_from_host(xs, gt)
    for _ in range(20):
        l1 = tr.step_from_host(xs, gt)
    assert np.isfinite(l0) and np.isfinite(l1) and l1 < l0


def test_submit_from_host_matches_blocking_steps():
    """The pipelined call (upload of step k+1 under the kernels of step k, loss read one step late) must walk
    the same trajectory as the blocking one: same losses, step by step, with changing batches."""
    from siren_mri_b200 import modules
    from siren_mri_b200.trainer import SirenTrainer
    n, steps = 2048, 12
    torch.manual_seed(3)
    xs = [torch.rand(1, n, 2).pin_memory() for _ in range(steps)]
    gts = [torch.rand(1, n, 1).pin_memory() for _ in range(steps)]

    def make():
        torch.manual_seed(7)
        m = modules.SingleBVPNet(in_features=2, out_features=1, precision="bf16").cuda()
        return SirenTrainer(m, n, lr=1e-4)

    tr_a = make()
    ref = [tr_a.step_from_host(x, g) for x, g in zip(xs, gts)]
    tr_b = make()
    got, prev = [], None
    for x, g in zip(xs, gts):
        h = tr_b.submit_from_host(x, g)
        if prev is not None:
            got.append(prev.result())
        prev = h
    got.append(prev.result())
    assert len(got) == steps
    for a, b in zip(ref, got):
        # same trajectory up to the order of the fp32 atomic accumulations in the gradient kernels
        assert abs(a - b) <= 1e-3 * abs(a), (ref, got)


def test_train_fast_matches_blocking_loop(tmp_path):
    """training.train_fast (sibling of the reference's training.train) walks the same trajectory as an explicit
    loop of blocking steps, writes the reference's checkpoint files, and its checkpoint loads back into a model."""
    import os
    from siren_mri_b200 import modules, training
    from siren_mri_b200.trainer import SirenTrainer
    n, epochs, per_epoch = 2048, 3, 4
    torch.manual_seed(5)
    data = [({"coords": torch.rand(1, n, 2) * 2 - 1}, {"img": torch.rand(1, n, 1)}) for _ in range(per_epoch)]

    def make():
        torch.manual_seed(11)
        return modules.SingleBVPNet(in_features=2, out_features=1, precision="bf16").cuda()

    m_a = make()
    tr = SirenTrainer(m_a, n, lr=1e-4, max_grad_norm=1.0)
    ref = [tr.step_from_host(mi["coords"].pin_memory(), g["img"].pin_memory()) for _ in range(epochs) for mi, g in data]
    m_b = make()
    seen = []
    losses = training.train_fast(m_b, data, epochs=epochs, lr=1e-4, steps_til_summary=5, epochs_til_checkpoint=2,
                                 model_dir=str(tmp_path / "run"), clip_grad=True, progress=lambda s: None,
                                 summary_fn=lambda m, mi, g, out, w, k: seen.append((k, tuple(out["model_out"].shape))))
    assert len(losses) == epochs * per_epoch
    for a, b in zip(ref, losses):
        assert abs(a - b) <= 1e-3 * abs(a), (ref, losses)
    assert seen == [(0, (1, n, 1)), (5, (1, n, 1)), (10, (1, n, 1))]
    ck = tmp_path / "run" / "checkpoints"
    assert sorted(os.listdir(ck)) == ["model_current.pth", "model_epoch_0002.pth", "model_final.pth",
                                      "train_losses_epoch_0002.txt", "train_losses_final.txt"]
    assert np.allclose(np.loadtxt(ck / "train_losses_final.txt"), losses)
    # the final checkpoint is the trained model: it loads into a fresh one and reproduces the blocking loop's weights
    m_c = make()
    m_c.load_state_dict(torch.load(ck / "model_final.pth"))
    x = data[0][0]["coords"].cuda()
    with torch.no_grad():
        y_a, y_c = m_a({"coords": x})["model_out"], m_c({"coords": x})["model_out"]
    assert rel_l2(y_c.cpu().numpy(), y_a.cpu().numpy()) < 1e-2


@pytest.mark.parametrize("clip", [0.0, 0.05])
def test_gradient_accumulation_matches_reference_loop(clip):
    """training.py:88-103 with accumulation_steps = 2: loss / 2 backward per batch, clip_grad_norm_ on the ACCUMULATED
    gradient after every batch, optimizer step + zero_grad after every second batch -- replayed with torch autograd
    on the composed model and compared with SirenTrainer.step(update=..., accumulation_steps=2)."""
    from siren_mri_b200 import modules
    from siren_mri_b200.trainer import SirenTrainer
    n, acc = 1500, 2
    torch.manual_seed(13)
    xs = [torch.rand(1, n, 2, device="cuda") * 2 - 1 for _ in range(4)]
    gts = [torch.rand(1, n, 1, device="cuda") * 2 - 1 for _ in range(4)]

    def make(backend):
        torch.manual_seed(17)
        return modules.SingleBVPNet(in_features=2, out_features=1, precision="fp32", backend=backend).cuda()

    ref = make("composed")
    opt = torch.optim.Adam(lr=1e-4, params=ref.parameters())
    ref_losses = []
    for k in range(4):
        out = ref({"coords": xs[k]})["model_out"]
        loss = ((out - gts[k]) ** 2).sum() / 16384.0
        ref_losses.append(float(loss))
        (loss / acc).backward()
        if clip:
            torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=clip)
        if (k + 1) % acc == 0:
            opt.step()
            opt.zero_grad()
    m = make("auto")
    tr = SirenTrainer(m, n, lr=1e-4, max_grad_norm=clip, precision="fp32")
    got = []
    for k in range(4):
        tr.coords.copy_(xs[k])
        tr.gt.copy_(gts[k])
        tr.step(update=(k + 1) % acc == 0, accumulation_steps=acc)
        got.append(float(tr.loss.item()) * acc)
    torch.cuda.synchronize()
    assert tr.steps == 2
    for a, b in zip(ref_losses, got):
        assert abs(a - b) < 2e-4 * abs(a), (ref_losses, got)
    for pr, pn in zip(ref.parameters(), m.parameters()):
        # an element whose gradient is below the mode's error may step the other way (+-lr): a handful per tensor
        du = (pn.detach() - pr.detach()).norm() / pr.detach().norm()
        assert float(du) < 1e-3, float(du)


@pytest.mark.parametrize("loss", ["sdf", "laplace_mse"])
def test_derivative_loss_tails_match_reference_loop(loss):
    """SirenTrainer(loss='sdf' | 'laplace_mse'): forward with jets, the loss tail kernel (siren_b200_sdf_grad /
    _laplace_mse_grad), reverse of the jets, clip, Adam in one graph -- against the reference's loop written out
    with torch autograd on the composed model: loss_functions.sdf (:460-484, clip_grad as train_sdf.py:57-60) /
    loss_functions.laplace_mse (:350-355), diff_operators.gradient / laplace, torch.optim.Adam."""
    from siren_mri_b200 import diff_operators, modules
    from siren_mri_b200.trainer import SirenTrainer
    from tools import workloads
    n = 3000
    d = 3 if loss == "sdf" else 2
    torch.manual_seed(3)
    x = torch.rand(1, n, d, device="cuda") * 2 - 1
    if loss == "sdf":
        sdf = torch.where(torch.arange(n, device="cuda") < n // 2, 0.0, -1.0).reshape(1, n, 1)
        normals = torch.nn.functional.normalize(torch.randn(1, n, 3, device="cuda"), dim=-1)
        normals = torch.where(sdf != -1, normals, -torch.ones_like(normals))
        gt = {"sdf": sdf, "normals": normals}
        clip = 1.0
    else:
        gt = {"laplace": 50.0 * torch.randn(1, n, 1, device="cuda")}
        clip = 0.0

    def make(backend):
        torch.manual_seed(19)
        return modules.SingleBVPNet(in_features=d, out_features=1, precision="fp32", backend=backend).cuda()

    ref = make("composed")
    opt = torch.optim.Adam(lr=1e-4, params=ref.parameters())
    ref_losses = []
    for _ in range(3):
        out = ref({"coords": x})
        if loss == "sdf":
            val = workloads.sdf_loss(out, gt, diff_operators.gradient)
        else:
            val = workloads.laplace_mse(out, gt, diff_operators.laplace)
        ref_losses.append(float(val.detach()))
        opt.zero_grad()
        val.backward()
        if clip:
            torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=clip)
        opt.step()
    m = make("auto")
    tr = SirenTrainer(m, n, lr=1e-4, max_grad_norm=clip, precision="fp32", loss=loss)
    got = [tr.step_from_host(x.cpu().pin_memory(), {k: v.cpu().pin_memory() for k, v in gt.items()}) for _ in range(3)]
    for a, b in zip(ref_losses, got):
        assert abs(a - b) < 5e-4 * abs(a), (ref_losses, got)
    for pr, pn in zip(ref.parameters(), m.parameters()):
        du = (pn.detach() - pr.detach()).norm() / pr.detach().norm()
        assert float(du) < 1e-3, float(du)
