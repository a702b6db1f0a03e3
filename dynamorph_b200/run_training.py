"""Drop-in for the VQ-VAE training path of /root/reference/run_training.py:
`run_one_batch` (:377-417) and `train` (:455-551).  The forward/backward run as the fused CUDA
schedule (dmb_train_forward / dmb_train_backward); `train` uses the fused flat Adam.
Dataset assembly, the ResNet loader path and the adversarial trainer are out of scope."""
from __future__ import annotations

import json
import os

import numpy as np
import torch as t

from .optim import FusedAdam
from .pipeline.train_utils import EarlyStopping


def draw_augmentation(n):
    """The reference's random draws, in its order (run_training.py:398,401): per sample a flip index in {0, 1, 2}
    then a rot90 count in {0..3}.  Returns one byte per sample, flip | rot << 2."""
    ops = np.empty(n, dtype=np.uint8)
    for i in range(n):
        flip_idx = int(np.random.choice([0, 1, 2]))
        rot_idx = int(np.random.choice([0, 1, 2, 3]))
        ops[i] = flip_idx | (rot_idx << 2)
    return ops


def augment_batch(batch):
    """run_training.py:396-403 -- per-sample flip over {none, H, W} then rot90 k in {0..3}; consumes
    np.random in the reference's order (two draws per sample).  CUDA batches of square patches are transformed by
    one kernel launch (dmb_augment_batch) instead of two tiny kernels per sample; the result is bit-identical."""
    if batch.is_cuda and batch.dim() == 4 and batch.shape[2] == batch.shape[3] and batch.dtype == t.float32:
        import ctypes as C
        from ._lib import call, ptr
        from .engine import _stream
        ops = t.from_numpy(draw_augmentation(len(batch))).to(batch.device, non_blocking=True)
        src = batch.contiguous()
        out = t.empty_like(src)
        B, Cc, H, W = src.shape
        call("dmb_augment_batch", ptr(src), ptr(ops), B, Cc, H, W, ptr(out), _stream())
        batch.copy_(out)            # the reference transforms `batch` in place
        return batch
    for idx_in_batch in range(len(batch)):
        img = batch[idx_in_batch]
        flip_idx = np.random.choice([0, 1, 2])
        if flip_idx != 0:
            img = t.flip(img, dims=(int(flip_idx),))
        rot_idx = int(np.random.choice([0, 1, 2, 3]))
        batch[idx_in_batch] = t.rot90(img, k=rot_idx, dims=[1, 2])
    return batch


def run_one_batch(model, batch, train_loss, model_kwargs=None, optimizer=None, transform=None, training=True):
    """Train (or validate) on a single batch; same contract as the reference, one host sync per call."""
    if transform is not None:
        batch = augment_batch(batch)
    _, train_loss_dict = model(batch, **(model_kwargs or {}))
    if training:
        train_loss_dict['total_loss'].backward()
        optimizer.step()
        model.zero_grad()
    keys = list(train_loss_dict.keys())
    vals = [train_loss_dict[k] for k in keys]
    tens = [v.detach().reshape(()) if isinstance(v, t.Tensor) else t.tensor(float(v), device=batch.device) for v in vals]
    host = t.stack([x.float() for x in tens]).tolist()          # the step's only device->host read
    for key, loss in zip(keys, host):
        train_loss.setdefault(key, []).append(float(loss))
    del batch, train_loss_dict
    return model, train_loss


def get_mask(mask, sample_ids, device='cuda:0'):
    """run_training.py:358-374 (second mask slice, rescaled to [0.5, 1])."""
    if mask is None:
        return None
    batch_mask = mask[sample_ids][0][:, 1:2, :, :]
    batch_mask = (batch_mask + 1.) / 2.
    return batch_mask.to(device)


class _ScalarLog:
    """TensorBoard when available (run_training.py:501,536-541), else a JSONL file with the same tags."""

    def __init__(self, output_dir):
        os.makedirs(output_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.w = SummaryWriter(output_dir)
        except Exception:
            self.w = None
            self.f = open(os.path.join(output_dir, "scalars.jsonl"), "a")

    def add_scalar(self, tag, value, step):
        if self.w is not None:
            self.w.add_scalar(tag, value, step)
        else:
            self.f.write(json.dumps({"tag": tag, "value": float(value), "step": int(step)}) + "\n")

    def flush(self):
        (self.w or self.f).flush()

    def close(self):
        (self.w or self.f).close()


def train(model, dataset, output_dir, relation_mat=None, mask=None, n_epochs=10, lr=0.001, batch_size=16,
          device='cuda:0', shuffle_data=False, transform=None, val_split_ratio=0.15, patience=20):
    """Legacy VAE trainer of the reference (run_training.py:455-551) on the fused step."""
    assert val_split_ratio is None or 0 < val_split_ratio < 1
    if patience is not None:
        assert val_split_ratio is not None
    if relation_mat is not None:
        raise NotImplementedError("time-matching loss is not part of the fused step yet (SURVEY.md section 8f, N3)")
    optimizer = FusedAdam(model, lr=lr, betas=(.9, .999))
    model.zero_grad()
    n_samples = len(dataset)
    sample_ids = list(range(n_samples))
    split = int(np.floor(val_split_ratio * n_samples))
    split_start = np.random.randint(0, n_samples - split)
    if shuffle_data:
        np.random.shuffle(sample_ids)
    val_ids = sample_ids[split_start: split_start + split]
    train_ids = sample_ids[:split_start] + sample_ids[split_start + split:]
    n_train, n_val = len(train_ids), len(val_ids)
    n_batches = int(np.ceil(n_train / batch_size))
    n_val_batches = int(np.ceil(n_val / batch_size))
    writer = _ScalarLog(output_dir)
    model_path = os.path.join(output_dir, 'model.pt')
    early_stopping = EarlyStopping(patience=patience, verbose=True, path=model_path)
    for epoch in range(n_epochs):
        train_loss, val_loss = {}, {}
        print('start epoch %d' % epoch)
        for i in range(n_batches):
            ids = train_ids[i * batch_size:min((i + 1) * batch_size, n_train)]
            batch = dataset[ids][0].to(device)
            model, train_loss = run_one_batch(model, batch, train_loss, optimizer=optimizer,
                                              model_kwargs={'time_matching_mat': None,
                                                            'batch_mask': get_mask(mask, ids, device)},
                                              transform=transform, training=True)
        for i in range(n_val_batches):
            ids = val_ids[i * batch_size:min((i + 1) * batch_size, n_val)]
            batch = dataset[ids][0].to(device)
            model, val_loss = run_one_batch(model, batch, val_loss, optimizer=optimizer,
                                            model_kwargs={'time_matching_mat': None,
                                                          'batch_mask': get_mask(mask, ids, device)},
                                            transform=transform, training=False)
        if shuffle_data:
            np.random.shuffle(train_ids)
        for key, loss in train_loss.items():
            train_loss[key] = sum(loss) / len(loss)
            writer.add_scalar('Loss/' + key, train_loss[key], epoch)
        for key, loss in val_loss.items():
            val_loss[key] = sum(loss) / len(loss)
            writer.add_scalar('Val loss/' + key, val_loss[key], epoch)
        early_stopping(val_loss['total_loss'], model)
        if early_stopping.early_stop:
            print("Early stopping")
            break
        writer.flush()
        print('epoch %d' % epoch)
        print('train: ', ''.join(['{}:{:0.4f}  '.format(key, loss) for key, loss in train_loss.items()]))
        print('validation: ', ''.join(['{}:{:0.4f}  '.format(key, loss) for key, loss in val_loss.items()]))
    writer.close()
    return model
