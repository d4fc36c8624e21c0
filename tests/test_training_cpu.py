"""Host logic of ``training.train_fast`` (sibling of the reference's training.train, training.py:19-146) with a
stand-in for the GPU step: directory layout, checkpoint cadence, loss bookkeeping (one step late, in order)."""
import os

import numpy as np
import pytest
import torch

from siren_mri_b200 import training


class _Handle:
    def __init__(self, v, log):
        self.v, self.log = v, log

    def result(self):
        self.log.append(("read", self.v))
        return self.v


class _FakeTrainer:
    """Loss of step k is k + mean(img); every submit nudges the model's first parameter (an 'update')."""

    def __init__(self, model, n_coords, lr, loss_weight, max_grad_norm):
        self.model, self.n, self.args = model, n_coords, (lr, loss_weight, max_grad_norm)
        self.k, self.log = 0, []

    def submit_from_host(self, coords, img, update=True, accumulation_steps=1):
        assert coords.dtype == torch.float32 and coords.is_contiguous() and coords.shape[-2] == self.n
        v = float(self.k + img.mean())
        self.log.append(("submit", v))
        self.updates = getattr(self, "updates", []) + [(bool(update), accumulation_steps)]
        with torch.no_grad():
            if update:
                next(self.model.parameters()).add_(1.0)
        self.k += 1
        return _Handle(v, self.log)


class _Net(torch.nn.Module):
    """dict in, dict out, like the reference's SingleBVPNet (modules.py:146-164)"""

    def __init__(self):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.zeros(1, 2))

    def forward(self, model_input):
        return {"model_in": model_input["coords"], "model_out": model_input["coords"] @ self.weight.t()}


def _loader(steps, n=16):
    return [({"coords": torch.zeros(1, n, 2)}, {"img": torch.full((1, n, 1), 0.5)}) for _ in range(steps)]


def test_train_fast_files_and_loss_order(tmp_path):
    model = _Net()
    made, seen, msgs = [], [], []

    def factory(*a):
        made.append(_FakeTrainer(*a))
        return made[-1]

    def summary_fn(m, mi, g, out, writer, step):
        seen.append((step, float(m.weight.detach()[0, 0]), tuple(out["model_out"].shape)))

    d = str(tmp_path / "run")
    losses = training.train_fast(model, _loader(3), epochs=4, lr=1e-4, steps_til_summary=5, epochs_til_checkpoint=2,
                                 model_dir=d, summary_fn=summary_fn, clip_grad=True, trainer_factory=factory,
                                 progress=msgs.append)
    assert len(made) == 1 and made[0].n == 16 and made[0].args == (1e-4, training.IMAGE_MSE_WEIGHT, 1.0)
    assert losses == [k + 0.5 for k in range(12)]                       # every step, in order
    ck = os.path.join(d, "checkpoints")
    assert sorted(os.listdir(ck)) == ["model_current.pth", "model_epoch_0002.pth", "model_final.pth",
                                      "train_losses_epoch_0002.txt", "train_losses_final.txt"]
    assert os.path.isdir(os.path.join(d, "summaries"))
    assert np.allclose(np.loadtxt(os.path.join(ck, "train_losses_final.txt")), losses)
    assert np.allclose(np.loadtxt(os.path.join(ck, "train_losses_epoch_0002.txt")), losses[:6])   # two epochs of three
    # checkpoints are state_dicts of the model; epoch 2 was saved after six updates, the final one after twelve
    assert float(torch.load(os.path.join(ck, "model_epoch_0002.pth"))["weight"][0, 0]) == 6.0
    assert float(torch.load(os.path.join(ck, "model_final.pth"))["weight"][0, 0]) == 12.0
    # summaries at steps 0, 5, 10 see the weights BEFORE that step's update, like the reference
    assert seen == [(0, 0.0, (1, 16, 1)), (5, 5.0, (1, 16, 1)), (10, 10.0, (1, 16, 1))]
    assert float(torch.load(os.path.join(ck, "model_current.pth"))["weight"][0, 0]) == 10.0
    assert len(msgs) == 3 and msgs[1].startswith("Epoch 1, Total loss 5.5")
    # pipelining: outside summary steps the loss of step k is read after step k+1 was submitted
    log = made[0].log
    assert log.index(("read", 1.5)) > log.index(("submit", 2.5))
    assert log.index(("read", 5.5)) < log.index(("submit", 6.5))       # summary step: read at once for the message


def test_train_fast_existing_dir(tmp_path):
    d = tmp_path / "run"
    d.mkdir()
    (d / "old.txt").write_text("x")
    model = _Net()
    with pytest.raises(FileExistsError):
        training.train_fast(model, _loader(1), 1, 1e-4, 10, 10, str(d), trainer_factory=_FakeTrainer)
    training.train_fast(model, _loader(1), 1, 1e-4, 10, 10, str(d), trainer_factory=_FakeTrainer, overwrite=True,
                        clip_grad=0.5, progress=lambda m: None)
    assert not (d / "old.txt").exists() and (d / "checkpoints" / "model_final.pth").exists()


def test_train_fast_gradient_accumulation_schedule(tmp_path):
    """training.py:99-103: the optimizer steps after every accumulation_steps-th batch and after an epoch's last."""
    model = _Net()
    made = []

    def factory(*a):
        made.append(_FakeTrainer(*a))
        return made[-1]

    losses = training.train_fast(model, _loader(5), epochs=2, lr=1e-4, steps_til_summary=100, epochs_til_checkpoint=10,
                                 model_dir=str(tmp_path / "acc"), trainer_factory=factory, progress=lambda m: None,
                                 accumulation_steps=2)
    assert len(losses) == 10
    assert made[0].updates == [(False, 2), (True, 2), (False, 2), (True, 2), (True, 2)] * 2
    assert float(model.weight.detach()[0, 0]) == 6.0


def test_train_fast_summary_sees_differentiable_output(tmp_path):
    """The reference's image summaries differentiate model_output w.r.t. model_in (utils.py write_image_summary)."""
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.ones(1, 2))

        def forward(self, model_input):
            x = model_input["coords"].clone().detach().requires_grad_(True)
            return {"model_in": x, "model_out": (x ** 2) @ self.weight.t()}

    got = []

    def summary_fn(m, mi, g, out, writer, step):
        grad = torch.autograd.grad(out["model_out"], out["model_in"], torch.ones_like(out["model_out"]), create_graph=True)[0]
        got.append(tuple(grad.shape))

    training.train_fast(Net(), _loader(2), epochs=1, lr=1e-4, steps_til_summary=1, epochs_til_checkpoint=10,
                        model_dir=str(tmp_path / "s"), trainer_factory=_FakeTrainer, progress=lambda m: None,
                        summary_fn=summary_fn)
    assert got == [(1, 16, 2), (1, 16, 2)]
