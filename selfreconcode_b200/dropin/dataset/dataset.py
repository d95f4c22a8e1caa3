"""Sequence dataset of the reference (dataset/dataset.py:9-237): same on-disk layout, attribute names and
accessors, so `train.py` / `infer.py` and `utils.save_model` / `load_model` see what they expect.

    <root>/imgs/%06d.(png|jpg)   BGR uint8, mapped to [-1,1]           (dataset.py:86-88)
    <root>/masks/%06d.png        foreground = any channel > 0          (dataset.py:97-98)
    <root>/normals/%06d.png      optional; RGB -> [-1,1]               (dataset.py:99-103)
    <root>/smpl_rec.npz          poses [F,24,3], trans [F,3], shape [10], gender, vid_seg_indices (optional)
    <root>/camera.npz            fx, fy, cx, cy, quat [4], T [3]

Per-frame learnables (poses, trans, latent code tables) live here as leaf tensors, exactly like the
reference: the DataLoader only carries images, `get_grad_parameters(ids)` slices the leaves so gradients reach
them.  Additions for multi-GPU runs: `ShardedSampler` (one disjoint, equally long index stream per rank,
SURVEY.md section 8e) and pinned host staging of the frame tensors.
"""
import os
import os.path as osp
import random
from glob import glob

import numpy as np
import torch

import utils


def _frame_id(path):
    return int(osp.basename(path).split('.')[0])


class SceneDataset(torch.utils.data.Dataset):
    def __init__(self, data_root, conds_lens={}, pin_memory=False):
        self.root = data_root
        self.pin_memory = bool(pin_memory) and torch.cuda.is_available()
        self.read_data()
        self.require_albedo = False
        self.conds, self.cond_ns = [], []
        for name, length in conds_lens.items():
            # latent codes start as smooth trajectories: random coefficients on the frame_num/5 lowest DCT modes
            k = max(self.frame_num // 5, 1)
            cond = (0.1 * torch.randn(length, k)).matmul(utils.DCTSpace(k, self.frame_num)).transpose(0, 1).contiguous()
            self.conds.append(cond.requires_grad_())
            self.cond_ns.append(name)

    def read_data(self):
        imgs = []
        for ext in ('.jpg', '.png'):
            imgs.extend(glob(osp.join(self.root, 'imgs/*' + ext)))
        imgs.sort(key=_frame_id)
        self.frame_num = len(imgs)
        self.img_ns = imgs
        self.mask_ns = []
        for i, name in enumerate(imgs):
            assert i == _frame_id(name), "frames must be numbered 0..F-1"
            m = osp.join(self.root, 'masks/%s.png' % osp.basename(name).split('.')[0])
            assert osp.isfile(m), m
            self.mask_ns.append(m)
        import cv2
        self.H, self.W, _ = cv2.imread(self.mask_ns[0]).shape
        d = np.load(osp.join(self.root, 'smpl_rec.npz'))
        self.poses = torch.from_numpy(d['poses'].astype(np.float32)).view(-1, 24, 3)
        self.trans = torch.from_numpy(d['trans'].astype(np.float32)).view(-1, 3)
        self.shape = torch.from_numpy(d['shape'].astype(np.float32)).view(-1)
        self.gender = str(d['gender']) if 'gender' in d else 'neutral'
        seg = d['vid_seg_indices'] if 'vid_seg_indices' in d else []
        self.video_segmented_index = list(np.asarray(seg).reshape(-1)[:-1].tolist()) if len(seg) else []
        c = np.load(osp.join(self.root, 'camera.npz'))
        f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(-1))
        self.camera_params = {'focal_length': f32([c['fx'], c['fy']]), 'princeple_points': f32([c['cx'], c['cy']]),
                              'cam2world_coord_quat': f32(c['quat']), 'world2cam_coord_trans': f32(c['T'])}

    def opt_camera_params(self, conf):
        keys = {'focal_length': 'focal_length', 'princeple_points': 'princeple_points',
                'cam2world_coord_quat': 'quat', 'world2cam_coord_trans': 'T'}
        for k, ck in keys.items():
            self.camera_params[k].requires_grad_(conf if isinstance(conf, bool) else conf.get_bool(ck))

    def learnable_weights(self):
        ws = [c for c in self.conds if c.requires_grad]
        ws += [v for v in self.camera_params.values() if v.requires_grad]
        ws += [v for v in (self.shape, self.poses, self.trans) if v.requires_grad]
        return ws

    def __len__(self):
        return self.frame_num

    def _stage(self, t):
        return t.pin_memory() if self.pin_memory else t

    def __getitem__(self, idx):
        import cv2
        out = {}
        img = cv2.imread(self.img_ns[idx]).astype(np.float32)
        out['img'] = self._stage(torch.from_numpy((img / 255. - 0.5) * 2).view(self.H, self.W, 3))
        mask = torch.from_numpy(cv2.imread(self.mask_ns[idx])) > 0
        out['mask'] = self._stage(mask.view(self.H, self.W, -1).any(-1).float())
        nf = self.img_ns[idx].replace('/imgs/', '/normals/')[:-3] + 'png'
        if osp.isfile(nf):
            out['normal'] = 2. * cv2.imread(nf)[:, :, ::-1].astype(np.float32) / 255. - 1.
        if self.require_albedo:
            alb = cv2.imread(osp.join(self.root, 'albedos/%d.png' % idx)).astype(np.float32)
            out['albedo'] = torch.from_numpy((alb / 255. - 0.5) * 2.).view(self.H, self.W, 3)
        return idx, out

    # the DataLoader cannot carry tensors that require grad: sliced here instead (dataset.py:116-122)
    def get_grad_parameters(self, idxs, device):
        conds = [c[idxs].to(device) for c in self.conds]
        if len(conds) < 2:
            conds = conds + [None] * (2 - len(conds))
        return (self.poses[idxs].to(device), self.trans[idxs].to(device), *conds)

    def get_camera_parameters(self, N, device):
        cp = self.camera_params
        return (cp['focal_length'].to(device).view(1, 2).expand(N, 2), cp['princeple_points'].to(device).view(1, 2).expand(N, 2),
                utils.quat2mat(cp['cam2world_coord_quat'].to(device).view(1, 4)).expand(N, 3, 3),
                cp['world2cam_coord_trans'].to(device).view(1, 3).expand(N, 3), self.H, self.W)

    def get_batchframe_data(self, name, fids, batchsize):
        """[len(fids), batchsize, ...] windows of consecutive frames centred on each id, shifted to stay inside
        the video (or inside the id's segment when the sequence is two videos) -> (windows, position of the id in
        its window)  (dataset.py:128-191)."""
        data = getattr(self, name)
        assert data.shape[0] >= self.frame_num
        data = data[:self.frame_num].to(fids.device)
        cuts = [0] + [int(c) for c in self.video_segmented_index] + [self.frame_num]
        if len(cuts) > 3:
            raise NotImplementedError("more than two video segments")
        starts = torch.full_like(fids, -1)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            assert batchsize < hi - lo
            sel = (fids >= lo) & (fids < hi)
            starts[sel] = (fids[sel] - batchsize // 2).clamp(min=lo, max=hi - batchsize)
        assert (starts >= 0).all().item()
        win = starts.view(-1, 1) + torch.arange(0, batchsize, device=fids.device).view(1, batchsize)
        return data[win], fids - starts


class ClipSampler(torch.utils.data.Sampler):
    """Whole clips of `clip_size` consecutive frames, clip order shuffled (dataset.py:195-215)."""

    def __init__(self, data_source, clip_size, shuffle):
        self.length, self.clip_size, self.shuffle = len(data_source), clip_size, shuffle
        self.n = self.length // clip_size
        if self.length == self.n * clip_size:
            self.n -= 1
        self.start = self.length - self.n * clip_size

    def __iter__(self):
        start = random.randrange(0, self.start + 1) if self.shuffle else 0
        out = torch.arange(start, start + self.n * self.clip_size).view(self.n, self.clip_size)
        if self.shuffle:
            out = out[torch.randperm(self.n)]
        return iter(out.view(-1).tolist())

    def __len__(self):
        return self.n * self.clip_size


class RandomSampler(torch.utils.data.Sampler):
    """Every `intersect`-th frame from a random phase, shuffled (dataset.py:217-237)."""

    def __init__(self, data_source, intersect, shuffle):
        self.length, self.intersect, self.shuffle = len(data_source), intersect, shuffle
        self.n = (self.length - 1) // intersect + 1
        self.start = self.length - intersect * (self.n - 1)

    def __iter__(self):
        if self.shuffle:
            index = torch.arange(random.randrange(0, self.start), self.length, self.intersect)
            index = index[torch.randperm(self.n)]
        else:
            index = torch.arange(0, self.length, self.intersect)
        assert index.numel() == self.n
        return iter(index.tolist())

    def __len__(self):
        return self.n


class ShardedSampler(torch.utils.data.Sampler):
    """Data-parallel split of an epoch: one shuffled permutation of the frames per epoch
    (same seed on every rank), padded to a multiple of the world size, rank r takes positions r, r+W, ...  All
    ranks get equally many frames, so per-frame means averaged across ranks equal the global mean."""

    def __init__(self, data_source, rank, world, shuffle=True, seed=0):
        self.length, self.rank, self.world, self.shuffle, self.seed = len(data_source), rank, world, shuffle, seed
        self.epoch = 0
        self.n = (self.length + world - 1) // world

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __iter__(self):
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            perm = torch.randperm(self.length, generator=g)
        else:
            perm = torch.arange(self.length)
        pad = self.n * self.world - self.length
        if pad:
            perm = torch.cat([perm, perm[:pad]])
        return iter(perm[self.rank::self.world].tolist())

    def __len__(self):
        return self.n


def getDatasetAndLoader(root, conds_lens, batch_size, shuffle, num_workers, opt_pose, opt_trans, opt_camera,
                        rank=0, world=1):
    dataset = SceneDataset(root, conds_lens)
    dataset.poses.requires_grad_(bool(opt_pose))
    dataset.trans.requires_grad_(bool(opt_trans))
    dataset.opt_camera_params(opt_camera)
    sampler = RandomSampler(dataset, 1, shuffle) if world == 1 else ShardedSampler(dataset, rank, world, shuffle)
    loader = torch.utils.data.DataLoader(dataset, batch_size, sampler=sampler, num_workers=num_workers)
    return dataset, loader


def write_sequence(root, imgs, masks, poses, trans, shape, camera, normals=None, gender='neutral'):
    """Writes a sequence in the layout above (used to build synthetic PeopleSnapshot-shaped sequences;
    the reference's own writer is people_snapshot_process.py:32-87).  imgs [F,H,W,3] in [-1,1] BGR,
    masks [F,H,W] {0,1}, normals [F,H,W,3] in [-1,1] RGB or None, camera = dict(fx,fy,cx,cy,quat,T)."""
    import cv2
    for sub in ('imgs', 'masks') + (('normals',) if normals is not None else ()):
        os.makedirs(osp.join(root, sub), exist_ok=True)
    to8 = lambda a: np.clip(np.round((np.asarray(a, dtype=np.float32) / 2. + 0.5) * 255.), 0, 255).astype(np.uint8)
    for i in range(len(imgs)):
        cv2.imwrite(osp.join(root, 'imgs/%06d.png' % i), to8(imgs[i]))
        cv2.imwrite(osp.join(root, 'masks/%06d.png' % i), (np.asarray(masks[i]) > 0).astype(np.uint8) * 255)
        if normals is not None:
            cv2.imwrite(osp.join(root, 'normals/%06d.png' % i), to8(normals[i])[:, :, ::-1])
    np.savez(osp.join(root, 'smpl_rec.npz'), poses=np.asarray(poses, np.float32), trans=np.asarray(trans, np.float32),
             shape=np.asarray(shape, np.float32), gender=gender)
    np.savez(osp.join(root, 'camera.npz'), **{k: np.asarray(v, np.float32) for k, v in camera.items()})
