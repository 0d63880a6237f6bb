"""kokoro-train (B200 path) end to end on a corpus-free synthetic dataset: two epochs with dynamic batching,
accumulation windows, SpecAugment switching on in the second epoch (a different CUDA graph per batch shape),
EMA validation and reference-named checkpoints."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cli_trains_two_epochs_on_synthetic_data(tmp_path, capsys):
    from kokoro_ruslan_b200 import cli
    rc = cli.main(["--synthetic", "14", "-e", "3", "-o", str(tmp_path), "--max-frames", "2400", "--min-batch-size", "1",
                   "--max-batch-size", "4", "--save-every", "1", "--val-split", "0.2", "--seed", "5"])
    assert rc == 0
    out = capsys.readouterr().out
    assert "epoch 3/3" in out and "done: 3 epochs" in out
    ck = torch.load(os.path.join(str(tmp_path), "checkpoint_epoch_3.pth"), weights_only=False)
    assert ck["epoch"] == 2 and ck["current_optimizer_step"] > 0
    assert len(ck["model_state_dict"]) == 311 and len(ck["ema_model_state_dict"]) == 311
    vals = [v for v in (ck["train_loss"], ck["val_loss"]) if v is not None]
    assert all(v == v and v < 100 for v in vals), vals          # finite losses
    w = ck["model_state_dict"]["mel_projection_out.weight"]
    assert torch.isfinite(w).all()
