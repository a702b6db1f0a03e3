"""Drop-in for the VQ-VAE training path of /root/reference/run_training.py: `run_one_batch` (:377-417),
`get_relation_tensor` (:335-355), `get_mask` (:358-374), `train` (:455-551), `main` (:771-948) and the
`-c cfg.yml` command line (:950-966).

`run_one_batch` keeps the reference's eager contract (autograd + any optimiser; forward / backward are the fused
CUDA schedule).  `train` does not go through it: the dataset lives on the GPU, every step is a CUDA-graph replay of
pack -> forward -> backward -> Adam (`trainer.FusedTrainer`), augmentation is one launch, the relation matrix of a
batch is densified on the device, and the per-batch losses stay on the device until the epoch ends (one read-back
per epoch instead of five per batch).  The host draws every random number in the reference's order (split start,
shuffles, two augmentation draws per sample for training AND validation batches), so a seeded run walks the same
batches as the reference.  ResNet / triplet / adversarial trainers are out of scope."""
from __future__ import annotations

import argparse
import json
import os
import pickle

import numpy as np
import torch as t

from ._lib import STREAM, call, ptr
from .pipeline.train_utils import EarlyStopping

LOSS_KEYS = ('recon_loss', 'commitment_loss', 'total_loss', 'perplexity', 'time_matching_loss')   # FusedTrainer order


# ----------------------------------------------------------------------------------------------- augmentation
def draw_augmentation(n):
    """The reference's random draws, in its order (run_training.py:398,401): per sample a flip index in {0, 1, 2}
    then a rot90 count in {0..3}.  Returns one byte per sample, flip | rot << 2."""
    ops = np.empty(n, dtype=np.uint8)
    for i in range(n):
        flip_idx = int(np.random.choice([0, 1, 2]))
        rot_idx = int(np.random.choice([0, 1, 2, 3]))
        ops[i] = flip_idx | (rot_idx << 2)
    return ops


def augment_batch(batch, out=None):
    """run_training.py:396-403 -- per-sample flip over {none, H, W} then rot90 k in {0..3} -- for a whole CUDA batch
    of square patches in ONE launch (dmb_augment_batch); consumes np.random exactly like the reference's loop (two
    draws per sample) and, like it, leaves the result in `batch`.  With `out` the result is written there instead
    (no copy back)."""
    if not (batch.is_cuda and batch.dim() == 4 and batch.shape[2] == batch.shape[3] and batch.dtype == t.float32):
        raise RuntimeError("dynamorph_b200: augmentation runs on the GPU: the batch must be a CUDA float32 tensor of "
                           f"square patches (B, C, S, S); got {tuple(batch.shape)} {batch.dtype} on {batch.device}")
    ops = t.from_numpy(draw_augmentation(len(batch))).to(batch.device, non_blocking=True)
    src = batch.contiguous()
    dst = t.empty_like(src) if out is None else out
    B, Cc, H, W = src.shape
    call("dmb_augment_batch", ptr(src), ptr(ops), B, Cc, H, W, ptr(dst), STREAM)
    if out is None:
        batch.copy_(dst)
        return batch
    return out


# ----------------------------------------------------------------------------------------------- eager step
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


def get_relation_tensor(relation_mat, sample_ids, device='cuda:0'):
    """Rows and columns `sample_ids` of the sparse pair-relation matrix as a dense float32 (B, B) tensor on `device`
    (run_training.py:335-355).  Only the batch's non-zeros cross PCIe; the dense block is assembled on the GPU."""
    if relation_mat is None:
        return None
    ids = np.asarray(sample_ids, dtype=np.int64)
    sub = relation_mat[ids, :][:, ids].tocoo()
    n = len(ids)
    if not device:
        return t.from_numpy(np.asarray(sub.todense())).float()
    dense = t.zeros(n, n, dtype=t.float32, device=device)
    if sub.nnz:
        flat = t.from_numpy(sub.row.astype(np.int64) * n + sub.col.astype(np.int64)).to(device, non_blocking=True)
        dense.view(-1).index_put_((flat,), t.from_numpy(sub.data.astype(np.float32)).to(device, non_blocking=True),
                                  accumulate=True)        # duplicate COO entries add up, like todense()
    return dense


def get_mask(mask, sample_ids, device='cuda:0'):
    """run_training.py:358-374 (second mask slice, rescaled to [0.5, 1])."""
    if mask is None:
        return None
    batch_mask = mask[sample_ids][0][:, 1:2, :, :]
    batch_mask = (batch_mask + 1.) / 2.
    return batch_mask.to(device)


# ----------------------------------------------------------------------------------------------- logging
class _ScalarLog:
    """Epoch scalars under the reference's tags (run_training.py:501,536-541): always as `scalars.jsonl` in the output
    directory, and through TensorBoard's SummaryWriter as well when that package is installed."""

    def __init__(self, output_dir):
        os.makedirs(output_dir, exist_ok=True)
        self.f = open(os.path.join(output_dir, "scalars.jsonl"), "a")
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.w = SummaryWriter(output_dir)
        except Exception:
            self.w = None

    def add_scalar(self, tag, value, step):
        self.f.write(json.dumps({"tag": tag, "value": float(value), "step": int(step)}) + "\n")
        if self.w is not None:
            self.w.add_scalar(tag, value, step)

    def flush(self):
        self.f.flush()
        if self.w is not None:
            self.w.flush()

    def close(self):
        self.f.close()
        if self.w is not None:
            self.w.close()


# ----------------------------------------------------------------------------------------------- train
class _ResidentSet:
    """The training tensor (and the mask tensor) on the GPU: a batch is one device gather.  Sets that do not fit in
    free HBM stay in pinned host memory and batches are gathered there and copied."""

    def __init__(self, tensor, device, reserve=0.5):
        tensor = tensor if tensor.dtype == t.float32 else tensor.float()
        need = tensor.numel() * 4
        free, _ = t.cuda.mem_get_info(device)
        self.device = device
        self.on_device = need < reserve * free
        self.data = tensor.to(device) if self.on_device else (tensor if tensor.is_pinned() else tensor.pin_memory())

    def take(self, ids):
        if self.on_device:
            return self.data.index_select(0, t.as_tensor(ids, dtype=t.int64).to(self.device, non_blocking=True))
        return self.data[t.as_tensor(ids, dtype=t.int64)].to(self.device, non_blocking=True)


def _slices(ids, batch_size):
    return [ids[a:a + batch_size] for a in range(0, len(ids), batch_size)]


def _epoch_means(dev_rows, n, total_last):
    """(n, 8) per-batch loss rows on the device -> {key: mean over batches}, the reference's `sum(loss)/len(loss)`
    of per-batch python floats (run_training.py:535,539) in the key order of the model's loss dict: ONE device->host
    read per phase per epoch."""
    if n == 0:
        return {}
    rows = dev_rows[:n].cpu().tolist()
    order = ('recon_loss', 'commitment_loss', 'time_matching_loss') + \
        (('perplexity', 'total_loss') if total_last else ('total_loss', 'perplexity'))
    return {k: sum(r[LOSS_KEYS.index(k)] for r in rows) / n for k in order}


def train(model, dataset, output_dir, relation_mat=None, mask=None, n_epochs=10, lr=0.001, batch_size=16,
          device='cuda:0', shuffle_data=False, transform=None, val_split_ratio=0.15, patience=20):
    """Legacy VAE trainer of the reference (run_training.py:455-551): same arguments, split, epoch structure,
    TensorBoard tags, early stopping and `model.pt` checkpoint; see the module docstring for how a step runs."""
    from .trainer import FusedTrainer
    assert val_split_ratio is None or 0 < val_split_ratio < 1
    if patience is not None:
        assert val_split_ratio is not None          # early stopping requires a validation set
    device = t.device(device)
    model = model.to(device).train()
    trainer = FusedTrainer(model, lr=lr, betas=(.9, .999))
    data = _ResidentSet(dataset.tensors[0], device)
    masks = None
    if mask is not None:                             # get_mask: second slice, rescaled to [0.5, 1]
        masks = _ResidentSet((mask.tensors[0][:, 1:2, :, :].float() + 1.) / 2., device)

    n_samples = len(dataset)
    order = list(range(n_samples))
    n_val = int(np.floor(val_split_ratio * n_samples)) if val_split_ratio is not None else 0
    val_at = np.random.randint(0, n_samples - n_val) if val_split_ratio is not None else 0
    if shuffle_data:
        np.random.shuffle(order)
    val_ids = order[val_at:val_at + n_val]
    train_ids = order[:val_at] + order[val_at + n_val:]

    log = _ScalarLog(output_dir)
    stopper = EarlyStopping(patience=patience, verbose=True, path=os.path.join(output_dir, 'model.pt'))
    staged = None
    for epoch in range(n_epochs):
        print('start epoch %d' % epoch)
        phases = (('Loss/', _slices(train_ids, batch_size), trainer.step),
                  ('Val loss/', _slices(val_ids, batch_size), trainer.forward_only))
        means = []
        for tag, batches, run in phases:
            rows = t.empty(max(len(batches), 1), 8, dtype=t.float32, device=device)
            for i, ids in enumerate(batches):
                batch = data.take(ids)
                if transform is not None:            # validation batches are augmented too (run_training.py:528-531)
                    if staged is None or staged.shape != batch.shape:
                        staged = t.empty_like(batch)
                    batch = augment_batch(batch, out=staged)
                rows[i].copy_(run(batch, batch_mask=None if masks is None else masks.take(ids),
                                  time_matching_mat=get_relation_tensor(relation_mat, ids, device=device)))
            means.append(_epoch_means(rows, len(batches), getattr(model, '_total_last', False)))
            for key, value in means[-1].items():
                log.add_scalar(tag + key, value, epoch)
        if shuffle_data:
            np.random.shuffle(train_ids)
        train_loss, val_loss = means
        if val_loss:
            stopper(val_loss['total_loss'], model)
            if stopper.early_stop:
                print("Early stopping")
                break
        log.flush()
        print('epoch %d' % epoch)
        print('train: ', ''.join('{}:{:0.4f}  '.format(k, v) for k, v in train_loss.items()))
        print('validation: ', ''.join('{}:{:0.4f}  '.format(k, v) for k, v in val_loss.items()))
    log.close()
    return model


# ----------------------------------------------------------------------------------------------- data assembly
def zscore(input_image, channel_mean=None, channel_std=None):
    """Dataset-wide per-channel z-score (pipeline/train_utils.py:228-250): given or estimated mean / std per channel,
    epsilon in the denominator.  Host-side preparation of the training set, done once."""
    x = np.asarray(input_image)
    mean = np.asarray(channel_mean if channel_mean else x.mean(axis=(0, 2, 3)), dtype=np.float64)
    std = np.asarray(channel_std if channel_std else x.std(axis=(0, 2, 3)), dtype=np.float64)
    print('channel_mean:', mean)
    print('channel_std:', std)
    shape = (1, -1) + (1,) * (x.ndim - 2)
    return (x - mean.reshape(shape)) / (std.reshape(shape) + np.finfo(float).eps)


def concat_relations(relations, labels, offsets):
    """Merge the pair-relation dicts / label arrays of several datasets, shifting sample ids by each dataset's
    offset (run_training.py:299-321)."""
    merged = {}
    for relation, offset in zip(relations, offsets):
        merged.update({(a + offset, b + offset): v for (a, b), v in relation.items()})
    return merged, np.concatenate([np.asarray(l) + o for l, o in zip(labels, offsets)], axis=0)


def reorder_with_trajectories(dataset, relations, seed=None):
    """Reorder the samples so that frames of one trajectory sit next to each other (run_training.py:97-159): repeatedly
    draw a random remaining sample; if it has adjacent-frame relations, emit its whole connected trajectory
    (breadth-first over the `== 2` pairs) in discovery order.  Returns (TensorDataset, csr relation matrix in the new
    order, the permutation).  The draws are the reference's (`np.random.choice(list(pool))`), so a seed reproduces
    its permutation."""
    from collections import deque
    from scipy.sparse import csr_matrix
    from torch.utils.data import TensorDataset
    if seed is not None:
        np.random.seed(seed)
    n = len(dataset)
    nxt = {}
    for (a, b), kind in relations.items():
        if kind == 2:
            nxt.setdefault(a, []).append(b)
    pool = set(range(n))
    perm = []
    while pool:
        start = np.random.choice(list(pool))
        group = [start]
        if start in nxt:
            todo = deque(group)
            while todo:
                for other in nxt[todo.popleft()]:
                    if other not in group:
                        group.append(other)
                        todo.append(other)
        perm.extend(group)
        pool.difference_update(group)
    perm = np.asarray(perm)
    pairs = np.asarray(list(relations.keys()), dtype=np.int64).reshape(-1, 2)
    kinds = np.asarray([v for v in relations.values()])
    keep = (kinds == 1) | (kinds == 2)
    # the reference appends every key but only the values 1 / 2, so any other value would misalign it; real relation
    # files hold nothing else
    assert keep.all(), "relations must be 1 (same trajectory) or 2 (adjacent frames)"
    mat = csr_matrix((kinds, (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    return TensorDataset(dataset.tensors[0][perm]), mat[perm][:, perm], list(perm)


def main(config_):
    """`python run_training.py -c cfg.yml` (run_training.py:771-948), VQ-VAE branch: load the assembled patches of every
    raw directory, z-score, merge relations, reorder by trajectory, build the `training.network` class and `train`."""
    from torch.utils.data import TensorDataset
    from .HiddenStateExtractor import vae
    from .configs.config_reader import YamlReader
    cfg = YamlReader().read_config(config_).training
    network = cfg.network
    if 'ResNet' in network:
        raise ValueError("network %s: the ResNet / triplet trainer is outside the VQ-VAE hot path (SURVEY.md section 2)"
                         % network)
    raw_dirs, train_dirs = cfg.raw_dirs, cfg.weights_dirs
    for d in train_dirs:
        os.makedirs(d, exist_ok=True)
    device = t.device('cuda:%d' % cfg.gpu_id)
    use_mask = getattr(cfg, 'use_mask', False)
    datasets, masks, relations, labels, offsets = [], [], [], [], [0]
    for raw_dir in raw_dirs:
        def load(name):
            path = os.path.join(raw_dir, name)
            print(f"\tloading {path}")
            with open(path, 'rb') as f:
                return pickle.load(f)
        ts_key = load('im_file_paths.pkl')
        patches = load('im_static_patches.pkl')
        print('dataset.shape:', patches.shape)
        labels.append(load('im_static_patches_labels.pkl'))
        relations.append(load('im_static_patches_relations.pkl'))      # order follows im_file_paths: do not sort
        print('len(ts_key):', len(ts_key))
        print('len(dataset):', len(patches))
        datasets.append(zscore(np.squeeze(patches), channel_mean=getattr(cfg, 'channel_mean', None),
                               channel_std=getattr(cfg, 'channel_std', None)).astype(np.float32))
        offsets.append(offsets[-1] + len(patches))
        if use_mask:
            masks.append(load('im_static_patches_mask.pkl'))
    dataset = TensorDataset(t.from_numpy(np.concatenate(datasets, axis=0)).float())
    relations, labels = concat_relations(relations, labels, offsets[:-1])
    dataset, relation_mat, order = reorder_with_trajectories(dataset, relations, seed=123)
    mask = None
    if use_mask:
        mask = TensorDataset(t.from_numpy(np.concatenate(masks, axis=0)[np.asarray(order)]).float())
    model = getattr(vae, network)(num_inputs=cfg.num_inputs,
                                  num_hiddens=cfg.num_hiddens,
                                  num_residual_hiddens=cfg.num_residual_hiddens,
                                  num_residual_layers=cfg.num_residual_layers,
                                  num_embeddings=cfg.num_embeddings,
                                  commitment_cost=cfg.commitment_cost,
                                  weight_matching=cfg.weight_matching,
                                  w_a=cfg.w_a, w_t=cfg.w_t, w_n=cfg.w_n, margin=cfg.margin,
                                  device=device).to(device)
    start = getattr(cfg, 'start_model_path', None)
    if start:
        print('Initialize the model with state {} ...'.format(start))
        model.load_state_dict(t.load(start, map_location=device))
    # the model is saved in the train directory of the last dataset (run_training.py:880)
    return train(model, dataset, output_dir=os.path.join(train_dirs[-1], cfg.model_name), relation_mat=relation_mat,
                 mask=mask, n_epochs=cfg.n_epochs, lr=cfg.learn_rate, batch_size=cfg.batch_size, device=device,
                 transform=True, val_split_ratio=cfg.val_split_ratio, patience=cfg.patience)


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('-c', '--config', type=str, required=True, help='path to yaml configuration file')
    return parser.parse_args(argv)


if __name__ == '__main__':
    main(parse_args().config)
