"""Device-resident keyframe store + mapping-window assembly.

What the reference does per mapped frame (src/Mapper.py:315-356, 528-541, 'global' keyframe selection):
  * a frame that becomes a keyframe is stored as a 10 % random pixel subset
    ``{gt_c2w, idx, color (P,3), depth (P,), est_c2w, rays_d (P,3)}`` (``torch.randperm(H*W)[:P]``);
  * every call of ``optimize_mapping`` re-stacks the subsets of ALL selected keyframes plus a fresh subset of the
    current frame into (K,P[,3]) tensors (``torch.stack`` of K list entries, ~1.1 GB of copies at K = 500), and every
    iteration ``get_samples_all`` rotates all K*P stored camera directions before gathering the few it needs.

Here the subsets live in preallocated (capacity,P[,3]) device tensors; a keyframe is appended by writing one row, the
mapping window is a *view* (no copy) whenever the selected frames are a contiguous run -- which is what the 'global'
policy selects outside a loop closure (all keyframes, Mapper.py:257-259) -- and the gather-then-rotate sampling is done
inside ``usl_ray_setup`` (mode 0) straight from these tensors.  The control policy around it (which frames to select on
a loop closure, keyframe_every) stays host code and is out of scope.

Pure torch plumbing, device-agnostic (the CPU tests compare it with the reference's own list-of-dicts bookkeeping).
"""
from typing import List, Optional, Sequence

import torch


class KeyframeStore:
    def __init__(self, capacity: int, H: int, W: int, device, pixel_ratio: float = 0.1):
        self.H, self.W = int(H), int(W)
        self.P = int(H * W * pixel_ratio)                       # num_pixels_to_save, Mapper.py:531
        self.capacity = int(capacity)
        f32 = dict(device=device, dtype=torch.float32)
        C, P = self.capacity + 1, self.P                        # +1: staging row for the current frame
        self.depth = torch.zeros((C, P), **f32)
        self.color = torch.zeros((C, P, 3), **f32)
        self.rays_d = torch.zeros((C, P, 3), **f32)             # camera-frame directions of the kept pixels
        self.est_c2w = torch.zeros((C, 4, 4), **f32)
        self.gt_c2w = torch.zeros((C, 4, 4), **f32)
        self.pixel_idx = torch.zeros((C, P), device=device, dtype=torch.int64)
        self.frame_idx: List[int] = []                          # keyframe_list (Mapper.py:526)
        self.device = device

    def __len__(self):
        return len(self.frame_idx)

    # ---- insertion -------------------------------------------------------------------------------------------------
    def _write(self, row, color_img, depth_img, dirs_cam, est_c2w, gt_c2w, indices):
        if indices is None:
            indices = torch.randperm(self.H * self.W)[:self.P]  # CPU generator, as the reference draws it (Mapper.py:532)
        ind = indices.to(self.device)
        if self.depth.is_cuda:
            # one launch gathers the three arrays into the store's row (usl_keyframe_insert)
            from . import _lib as L
            f32 = lambda t: t.to(self.device, torch.float32).contiguous()
            L.call("usl_keyframe_insert", L.ptr(f32(color_img)), L.ptr(f32(depth_img)), L.ptr(f32(dirs_cam)), L.ptr(ind.to(torch.int64).contiguous()),
                   self.P, L.ptr(self.color[row]), L.ptr(self.depth[row]), L.ptr(self.rays_d[row]), L.ptr(self.pixel_idx[row]), L.stream())
            self.est_c2w[row] = est_c2w
            if gt_c2w is not None:
                self.gt_c2w[row] = gt_c2w
            return indices
        self.color[row] = color_img.reshape(-1, 3)[ind]
        self.depth[row] = depth_img.reshape(-1)[ind]
        self.rays_d[row] = dirs_cam.reshape(-1, 3)[ind]
        self.est_c2w[row] = est_c2w
        if gt_c2w is not None:
            self.gt_c2w[row] = gt_c2w
        self.pixel_idx[row] = ind
        return indices

    def stage_current(self, color_img, depth_img, dirs_cam, cur_c2w, gt_c2w=None, indices: Optional[torch.Tensor] = None):
        """The current frame's subset for this call of optimize_mapping (Mapper.py:330-344): written to the row after
        the last keyframe, so the window [keyframes..., current] stays one contiguous view."""
        return self._write(len(self), color_img, depth_img, dirs_cam, cur_c2w, gt_c2w, indices)

    def append(self, idx: int, color_img, depth_img, dirs_cam, est_c2w, gt_c2w=None, indices: Optional[torch.Tensor] = None):
        """Add frame ``idx`` as a keyframe with a freshly drawn subset (Mapper.py:526-541)."""
        if len(self) >= self.capacity:
            raise RuntimeError(f"KeyframeStore: capacity {self.capacity} exhausted")
        indices = self._write(len(self), color_img, depth_img, dirs_cam, est_c2w, gt_c2w, indices)
        self.frame_idx.append(int(idx))
        return indices

    def promote_staged(self, idx: int, est_c2w=None):
        """Keep the staged subset of the current frame as the new keyframe (no second randperm / gather: the fused
        driver's choice; the reference redraws, use append() for its exact RNG consumption)."""
        if len(self) >= self.capacity:
            raise RuntimeError(f"KeyframeStore: capacity {self.capacity} exhausted")
        if est_c2w is not None:
            self.est_c2w[len(self)] = est_c2w
        self.frame_idx.append(int(idx))

    # ---- window assembly ---------------------------------------------------------------------------------------------
    def window(self, frames: Optional[Sequence[int]] = None, with_current: bool = True):
        """(c2ws, depths, colors, rays_d) of the selected keyframes (+ the staged current frame last), i.e. the stacked
        tensors of Mapper.py:346-351.  frames=None selects every keyframe.  Views when the selection is a contiguous
        run ending at the last keyframe, gathered copies otherwise."""
        K = len(self)
        frames = list(range(K)) if frames is None else [int(f) for f in frames]
        rows = frames + ([K] if with_current else [])
        if rows and rows == list(range(rows[0], rows[0] + len(rows))):
            sl = slice(rows[0], rows[0] + len(rows))
            return self.est_c2w[sl], self.depth[sl], self.color[sl], self.rays_d[sl]
        r = torch.tensor(rows, device=self.device, dtype=torch.int64)
        return self.est_c2w[r], self.depth[r], self.color[r], self.rays_d[r]

    def mapping_batches(self, idx_main, n_main: int, idx_recent=None, n_recent: int = 0, frames=None, c2ws=None):
        """Batches in the layout MappingStep.run takes: the main get_samples_all call over the whole window and,
        when given, the extra 200 px x last-10-frames call (Mapper.py:379-393).  c2ws overrides the stored poses."""
        cw, dep, col, rd = self.window(frames)
        cw = cw if c2ws is None else c2ws
        Kw = dep.shape[0]
        out = [(cw, dep, col, rd, idx_main, n_main, 0)]
        if idx_recent is not None and n_recent > 0:
            out.append((cw[Kw - 10:], dep[Kw - 10:], col[Kw - 10:], rd[Kw - 10:], idx_recent, n_recent, Kw - 10))
        return out

    def write_back_poses(self, c2ws: torch.Tensor, frames: Optional[Sequence[int]] = None):
        """Put jointly optimised poses back (Mapper.py:447-457): c2ws holds frames[1:] + [current] (the first window
        frame is fixed). Returns the optimised pose of the current frame."""
        K = len(self)
        frames = list(range(K)) if frames is None else [int(f) for f in frames]
        for j, f in enumerate(frames[1:]):
            self.est_c2w[f] = c2ws[j]
        self.est_c2w[K] = c2ws[-1]
        return c2ws[-1]

    # ---- co-visibility (keyframe_selection_LC, Mapper.py:177-236) -----------------------------------------------------
    def covisibility(self, rays_o, rays_d, gt_depth, H, W, fx, fy, cx, cy, num_samples: int = 8, edge: float = 20.0, skip_last: int = 2):
        """percent_inside over the stored keyframes except the last `skip_last` (which the mapper always includes,
        Mapper.py:214): one launch, result stays on the device.  The selection policy built on it stays host code."""
        from . import _lib as L
        K = len(self) - skip_last
        out = torch.zeros((max(K, 0),), device=self.device, dtype=torch.float32)
        if K > 0:
            L.call("usl_keyframe_covisibility", L.ptr(rays_o.contiguous()), L.ptr(rays_d.contiguous()), L.ptr(gt_depth.contiguous()),
                   rays_o.shape[0], int(num_samples), L.ptr(self.est_c2w[:K].contiguous()), K, int(H), int(W), float(fx), float(fy), float(cx),
                   float(cy), float(edge), L.ptr(out), L.stream())
        return out

    # ---- wire format ---------------------------------------------------------------------------------------------------
    def as_dicts(self):
        """The reference's keyframe_dict list (Mapper.py:540-541), as views into the store (Mesher / Logger consumers)."""
        return [{"gt_c2w": self.gt_c2w[k], "idx": self.frame_idx[k], "color": self.color[k], "depth": self.depth[k],
                 "est_c2w": self.est_c2w[k], "rays_d": self.rays_d[k]} for k in range(len(self))]
