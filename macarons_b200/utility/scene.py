"""Scene / Cell: the regular cell grid that stores the reconstructed surface points and the proxy points of a MACARONS
scene (reference macarons/utility/macarons_utils.py:2503-2932), mirrored as far as the NBV scoring path uses it
(SURVEY.md section 8f rank 2): filling cells with depth-map points, cell look-ups, the proxy-point state tensors and
their per-frame updates.  Same class / method / attribute names and the same results (incl. the order in which the
global CPU generator is consumed by the capacity sub-sampling) so that the reference's testers can use either.

Layout: all per-proxy-point state lives in flat device tensors (`proxy_points`, `view_states`, `proxy_proba`, ...);
a cell is a small record (box, resolution, capacity) plus its (n, 3) points and (n, feature_dim) features.  The view
state update runs the binning kernel (csrc/viewstate.cu through `compute_view_state`); everything else here is
bookkeeping on index tensors."""
import numpy as np
import torch

from .scone_utils import compute_view_state


def _float_floor_divide(x, d):
    """reference utility/utils.py:113-117: float floor division built on torch.remainder"""
    return (x - x % d) / d


def _largest_face_diagonal_area(l, w, h):
    return max((l * torch.sqrt(w ** 2 + h ** 2)).item(), (w * torch.sqrt(h ** 2 + l ** 2)).item(),
               (h * torch.sqrt(l ** 2 + w ** 2)).item())


class Cell:
    """One box of the grid (reference :2503-2583).  `capacity` points at most, no two closer than `resolution`;
    one of the two may be None and is then derived from the other through the area of the largest diagonal section."""

    def __init__(self, center, l, w, h, capacity, resolution, device, feature_dim=0):
        self.center, self.l, self.w, self.h, self.device = center, l, w, h, device
        half = torch.Tensor([[l / 2., w / 2., h / 2.]]).to(center.device)
        self.x_min, self.x_max = center - half, center + half
        if resolution is None and capacity is None:
            raise NameError("Please choose a capacity or a resolution.")
        if resolution is None:
            area_per_point = _largest_face_diagonal_area(l, w, h) / capacity
            resolution = 2 * np.sqrt(area_per_point / np.pi)
        elif capacity is None:
            capacity = int(_largest_face_diagonal_area(l, w, h) // (np.pi * (resolution / 2.) ** 2))
        self.capacity, self.resolution = capacity, resolution
        self.use_feature = feature_dim > 0
        if self.use_feature:
            self.feature_dim = feature_dim
        self.empty()

    def empty(self):
        self.cell_pts = torch.zeros(0, 3).to(self.device)
        if self.use_feature:
            self.cell_features = torch.zeros(0, self.feature_dim).to(self.device)

    def is_empty(self):
        return self.cell_pts.shape[0] == 0

    def fill(self, pts, features=None, n_point_min=0):
        """Add the points strictly inside the box that are further than `resolution` from every stored point, then keep
        a random subset of `capacity` points (torch.randperm on the global CPU generator, like the reference :2566)."""
        with_features = self.use_feature and features is not None
        inside = (torch.max(pts - self.x_max, dim=-1)[0] < 0.) & (torch.min(pts - self.x_min, dim=-1)[0] > 0.)
        new_pts = pts[inside]
        if new_pts.shape[0] <= n_point_min:
            return
        new_features = features[inside] if with_features else None
        if self.cell_pts.shape[0] > 0:
            far = torch.min(torch.cdist(new_pts.double(), self.cell_pts.double(), p=2.0), dim=-1)[0] > self.resolution
            new_pts = new_pts[far]
            if with_features:
                new_features = new_features[far]
        self.cell_pts = torch.vstack((self.cell_pts, new_pts))
        keep = torch.randperm(len(self.cell_pts))[:self.capacity]
        self.cell_pts = self.cell_pts[keep]
        if with_features:
            self.cell_features = torch.vstack((self.cell_features, new_features))[keep]


class Scene:
    """grid_l x grid_w x grid_h cells over the box [x_min, x_max] (reference :2586-2932)."""

    def __init__(self, x_min, x_max, grid_l, grid_w, grid_h, cell_capacity, cell_resolution, n_proxy_points, device,
                 view_state_n_elev=7, view_state_n_azim=2 * 7, feature_dim=0, mirrored_scene=False, score_threshold=1.,
                 mirrored_axis=None):
        self.grid_l, self.grid_w, self.grid_h = grid_l, grid_w, grid_h
        self.x_min, self.x_max = 0. + x_min, 0. + x_max
        self.mirrored_scene, self.mirrored_axis = mirrored_scene, mirrored_axis
        if mirrored_scene:
            if mirrored_axis is None:
                raise NameError("Please provide the list of mirrored axis.")
            for axis in mirrored_axis:
                self.x_min[..., axis], self.x_max[..., axis] = -self.x_max[..., axis], -self.x_min[..., axis]
        size = self.x_max - self.x_min
        self.l, self.w, self.h = size[0] / grid_l, size[1] / grid_w, size[2] / grid_h

        self.cells = {}
        for i_l in range(grid_l):
            for i_w in range(grid_w):
                for i_h in range(grid_h):
                    center = torch.Tensor([self.x_min[0] + (1 / 2. + i_l) * self.l, self.x_min[1] + (1 / 2. + i_w) * self.w,
                                           self.x_min[2] + (1 / 2. + i_h) * self.h]).to(device)
                    cell = Cell(center=center, l=self.l, w=self.w, h=self.h, capacity=cell_capacity,
                                resolution=cell_resolution, device=device, feature_dim=feature_dim)
                    self.cells[str([i_l, i_w, i_h])] = cell
                    # the first cell fixes whichever of the two was left open (reference :2626-2629)
                    cell_resolution, cell_capacity = cell.resolution, cell.capacity
        self.cell_resolution, self.cell_capacity = cell_resolution, cell_capacity
        self.feature_dim, self.device = feature_dim, device
        # per-axis box tables of the grid (the cells' own fp32 x_min / x_max values) for the batched fill
        pick = lambda attr, axis, n, key: torch.stack([getattr(self.cells[str(key(i))], attr)[0, axis] for i in range(n)])
        self._axis_min = [pick("x_min", 0, grid_l, lambda i: [i, 0, 0]), pick("x_min", 1, grid_w, lambda i: [0, i, 0]),
                          pick("x_min", 2, grid_h, lambda i: [0, 0, i])]
        self._axis_max = [pick("x_max", 0, grid_l, lambda i: [i, 0, 0]), pick("x_max", 1, grid_w, lambda i: [0, i, 0]),
                          pick("x_max", 2, grid_h, lambda i: [0, 0, i])]
        self.batched_fill = True      # False: fill cell by cell like the reference (the two give identical states)

        # proxy points: call initialize_proxy_points() before use
        self.n_proxy_points = n_proxy_points
        self.proxy_points = self.proxy_proba = self.proxy_supervision_occ = self.view_states = None
        self.view_state_n_elev, self.view_state_n_azim = view_state_n_elev, view_state_n_azim
        self.n_view_state_cameras = view_state_n_azim * view_state_n_elev
        self.proxy_n_inside_fov = self.proxy_n_behind_depth = self.out_of_field = None
        self.score_threshold = score_threshold
        # typical spacing of the proxy points: diameter of a ball of the volume each point stands for
        volume_per_proxy_point = self.l * self.w * self.h / (n_proxy_points / (grid_l * grid_h * grid_w))
        self.distance_between_proxy_points = 2 * np.power(3 * volume_per_proxy_point.item() / (4 * np.pi), 1. / 3.)

    # ---- geometry look-ups ---------------------------------------------------------------------------------------
    def get_pts_in_bounding_box(self, pts, return_mask=True):
        pts_mask = ((pts >= self.x_min) * (pts <= self.x_max)).prod(dim=-1).bool()
        return (pts[pts_mask], pts_mask) if return_mask else pts[pts_mask]

    def get_cells_for_each_pt(self, pts):
        """(n, 3) -> (n, 3) int64 cell index of every point, clamped to the grid (reference :2684-2698)."""
        rel = pts - self.x_min
        i_l = _float_floor_divide(rel[:, 0:1], self.l)
        i_w = _float_floor_divide(rel[:, 1:2], self.w)
        i_h = _float_floor_divide(rel[:, 2:3], self.h)
        i_l[i_l >= self.grid_l] = self.grid_l - 1
        i_w[i_w >= self.grid_w] = self.grid_w - 1
        i_h[i_h >= self.grid_h] = self.grid_h - 1
        res = torch.hstack((i_l, i_w, i_h)).long()
        res[res < 0] = 0
        return res

    def get_englobing_cells(self, pts, list=False):
        res = torch.unique(self.get_cells_for_each_pt(pts), dim=0)
        return res.cpu().numpy().tolist() if list else res

    def get_neighboring_cells(self, cell_idx):
        """The (up to 27) cells around `cell_idx`, itself included, clamped to the grid, sorted and unique."""
        shift = torch.cartesian_prod(torch.arange(0, 3), torch.arange(0, 3), torch.arange(0, 3)).to(self.device) - 1
        res = cell_idx + shift
        hi = torch.tensor([self.grid_l - 1, self.grid_w - 1, self.grid_h - 1], device=res.device)
        res = torch.minimum(torch.clamp(res, min=0), hi)
        return torch.unique(res, dim=0)

    def get_key_from_idx(self, cell_idx):
        return str(cell_idx.cpu().numpy().tolist())

    # ---- cell contents -------------------------------------------------------------------------------------------
    def fill_cells(self, pts, features=None, n_point_min=0):
        pts_inside, inside_mask = self.get_pts_in_bounding_box(pts, return_mask=True)
        fts_inside = features[inside_mask] if features is not None else None
        cell_list = self.get_englobing_cells(pts_inside, list=True)
        if self.batched_fill and len(cell_list) > 1 and self._fill_cells_batched(pts_inside, fts_inside, n_point_min, cell_list):
            return
        for cell_idx in cell_list:
            self.cells[str(cell_idx)].fill(pts_inside, features=fts_inside, n_point_min=n_point_min)

    def _fill_cells_batched(self, pts, features, n_point_min, cell_list):
        """`for cell in cell_list: cell.fill(pts, features, n_point_min)` for all cells at once, with the same result and
        the same consumption of the global CPU generator (one torch.randperm per filled cell, in list order):
        every point is assigned to the cell whose open box contains it (the cells' own fp32 bounds, tested for the
        floor-division cell and its neighbours along every axis), the points are sorted by cell, ONE kernel computes the
        distance of every new point to the stored points of its own cell (the per-cell float64 cdist of the reference),
        and one gather applies all capacity sub-samplings.  ~20 tensor ops and 3 host synchronisations per call instead
        of ~15 ops and 3 synchronisations per cell.  Returns False (nothing changed) in the rare case that rounding puts
        a point inside two neighbouring boxes; the caller then fills cell by cell."""
        from .. import ops
        dev = pts.device
        grid = (self.grid_l, self.grid_w, self.grid_h)
        with_features = features is not None and self.feature_dim > 0
        floor = self.get_cells_for_each_pt(pts)                                     # (M, 3) int64
        index = []
        ok = torch.ones(pts.shape[0], dtype=torch.bool, device=dev)
        ambiguous = torch.zeros((), dtype=torch.bool, device=dev)
        for a in range(3):
            lo_tab, hi_tab = self._axis_min[a].to(dev), self._axis_max[a].to(dev)
            chosen = torch.full((pts.shape[0],), -1, dtype=torch.int64, device=dev)
            hits = torch.zeros(pts.shape[0], dtype=torch.int64, device=dev)
            for o in (-1, 0, 1):
                c = floor[:, a] + o
                valid = (c >= 0) & (c < grid[a])
                cc = c.clamp(0, grid[a] - 1)
                inside = valid & (pts[:, a] - hi_tab[cc] < 0.) & (pts[:, a] - lo_tab[cc] > 0.)
                chosen = torch.where(inside, cc, chosen)
                hits = hits + inside.long()
            ambiguous = ambiguous | (hits > 1).any()
            ok = ok & (hits == 1)
            index.append(chosen)
        lin = (index[0] * grid[1] + index[1]) * grid[2] + index[2]
        n_cells = grid[0] * grid[1] * grid[2]
        listed = torch.zeros(n_cells, dtype=torch.bool, device=dev)
        list_ids = [(c[0] * grid[1] + c[1]) * grid[2] + c[2] for c in cell_list]
        listed[torch.tensor(list_ids, device=dev)] = True
        ok = ok & listed[lin.clamp(0, n_cells - 1)]
        lin = torch.where(ok, lin, torch.full_like(lin, n_cells))                   # unassigned points sort to the end
        order = torch.argsort(lin, stable=True)
        counts_dev = torch.bincount(lin, minlength=n_cells + 1)
        host = torch.cat((counts_dev, ambiguous.long().view(1))).cpu()              # synchronisation 1
        if int(host[-1]) != 0:
            return False
        counts = host[:n_cells].tolist()
        work = [(cid, self.cells[str(c)]) for cid, c in zip(list_ids, cell_list) if counts[cid] > n_point_min]
        if not work:
            return True
        starts = torch.cumsum(host[:n_cells], 0) - host[:n_cells]
        # new points of the cells to fill, cell by cell in list order (inside a cell: original order, as pts[mask] gives)
        take = torch.cat([order[int(starts[cid]):int(starts[cid]) + counts[cid]] for cid, _ in work]) if len(work) > 1 \
            else order[int(starts[work[0][0]]):int(starts[work[0][0]]) + counts[work[0][0]]]
        new_pts = pts[take]
        new_fts = features[take] if with_features else None
        n_new = [counts[cid] for cid, _ in work]
        n_old = [int(cell.cell_pts.shape[0]) for _, cell in work]
        slot = torch.repeat_interleave(torch.arange(len(work), device=dev), torch.tensor(n_new, device=dev))
        old_pts = torch.cat([cell.cell_pts for _, cell in work])
        if sum(n_old) > 0:
            off = torch.tensor([0] + list(torch.tensor(n_old).cumsum(0).tolist()), dtype=torch.int32, device=dev)
            if new_pts.is_cuda:
                dist = ops.cell_min_dist(new_pts.to(torch.float32), slot.to(torch.int32), old_pts.to(torch.float32), off)
            else:       # host tensors (tests without a GPU): the reference's per-cell cdist
                dist = torch.full((new_pts.shape[0],), float("inf"), dtype=torch.float64)
                a0 = 0
                for s_i, (n_n, n_o) in enumerate(zip(n_new, n_old)):
                    if n_o:
                        o0 = int(off[s_i])
                        dist[a0:a0 + n_n] = torch.min(torch.cdist(new_pts[a0:a0 + n_n].double(), old_pts[o0:o0 + n_o].double(),
                                                                  p=2.0), dim=-1)[0]
                    a0 += n_n
            resolution = torch.tensor([cell.resolution for _, cell in work], dtype=torch.float64, device=dev)[slot]
            keep = dist > resolution
            kept = torch.nonzero(keep).view(-1)                                     # synchronisation 2
            n_keep = torch.bincount(slot[kept], minlength=len(work)).tolist()       # synchronisation 3
            new_pts = new_pts[kept]
            if with_features:
                new_fts = new_fts[kept]
        else:
            n_keep = n_new
        # capacity sub-sampling of every cell: randperm over [stored | kept new], drawn in list order
        gather, sizes = [], []
        old_off, new_off, E = 0, 0, sum(n_old)
        for (cid, cell), n_o, n_k in zip(work, n_old, n_keep):
            perm = torch.randperm(n_o + n_k)[:cell.capacity]
            gather.append(torch.where(perm < n_o, perm + old_off, perm - n_o + E + new_off))
            sizes.append(perm.numel())
            old_off, new_off = old_off + n_o, new_off + n_k
        gidx = torch.cat(gather).to(dev)
        all_pts = torch.cat((old_pts, new_pts)).index_select(0, gidx)
        pieces = torch.split(all_pts, sizes)
        if with_features:
            old_fts = torch.cat([cell.cell_features for _, cell in work])
            f_pieces = torch.split(torch.cat((old_fts, new_fts.to(old_fts.dtype))).index_select(0, gidx), sizes)
        for i, (_, cell) in enumerate(work):
            cell.cell_pts = pieces[i]
            if with_features:
                cell.cell_features = f_pieces[i]
        return True

    def empty_cells(self):
        for cell in self.cells.values():
            cell.empty()

    def get_pt_cloud_from_cells(self, cell_indices, return_features=True):
        with_features = return_features and (self.feature_dim > 0)
        keys = [str(c) for c in cell_indices.cpu().numpy().tolist()] if len(cell_indices.shape) > 1 \
            else [str(cell_indices.cpu().numpy().tolist())]
        cells = [self.cells[k] for k in keys]
        pts = torch.vstack([torch.zeros(0, 3, device=self.device)] + [c.cell_pts for c in cells])
        if with_features:
            return pts, torch.vstack([torch.zeros(0, self.feature_dim, device=self.device)] + [c.cell_features for c in cells])
        return pts

    def return_entire_pt_cloud(self, return_features=True):
        with_features = return_features and (self.feature_dim > 0)
        cells = list(self.cells.values())
        pts = torch.vstack([torch.zeros(0, 3, device=self.device)] + [c.cell_pts for c in cells])
        if with_features:
            return pts, torch.vstack([torch.zeros(0, self.feature_dim, device=self.device)] + [c.cell_features for c in cells])
        return pts

    def set_all_features_to_value(self, value):
        for cell in self.cells.values():
            if len(cell.cell_features) > 0:
                cell.cell_features = torch.full_like(cell.cell_features, value)

    # ---- proxy points --------------------------------------------------------------------------------------------
    def sample_in_box(self, n_sample):
        return self.x_min + (self.x_max - self.x_min) * torch.rand(n_sample, 3, device=self.device)

    def initialize_proxy_points(self, n_proxy_points=None, default_proba_value=0.5):
        n = self.n_proxy_points if n_proxy_points is None else n_proxy_points
        self.proxy_points = self.sample_in_box(n)
        self.proxy_proba = torch.zeros(n, 1, device=self.device) + default_proba_value
        self.proxy_supervision_occ = torch.ones(n, 1, device=self.device)
        self.view_states = torch.zeros(n, self.n_view_state_cameras, device=self.device)
        self.out_of_field = torch.ones(n, 1, device=self.device)
        self.proxy_n_inside_fov = torch.zeros(n, 1, device=self.device)
        self.proxy_n_behind_depth = torch.zeros(n, 1, device=self.device)

    def get_proxy_indices_from_mask(self, proxy_mask):
        return torch.arange(start=0, end=self.n_proxy_points, device=self.device).view(-1, 1)[proxy_mask]

    def get_proxy_mask_from_indices(self, proxy_indices):
        mask = torch.zeros(self.n_proxy_points, device=self.device).bool()
        mask[proxy_indices.view(-1).long()] = True
        return mask

    def update_proxy_view_states(self, camera, proxy_mask, signed_distances=None, distance_to_surface=None, X_cam=None):
        """OR the bin of the current camera position into the view state of the masked proxy points (world space);
        with `signed_distances` only the points closer than `distance_to_surface` in front of / behind the depth map
        are updated (reference :2818-2877).  The binning is the CUDA kernel behind compute_view_state."""
        if signed_distances is not None:
            if distance_to_surface is None:
                distance_to_surface = 3 * self.distance_between_proxy_points
            update_mask = torch.zeros_like(proxy_mask).bool()
            update_mask[proxy_mask] = signed_distances.view(-1) < distance_to_surface
        else:
            update_mask = proxy_mask
        pts = self.proxy_points[update_mask]
        if X_cam is None:
            X_cam = camera.X_cam
        if pts.shape[0] == 0:
            return
        seen = compute_view_state(pts.view(1, -1, 3), X_cam, self.view_state_n_elev, self.view_state_n_azim)
        # accumulated states stay in {0, 1}: heaviside(old + new, 0) == 1 where either is set
        self.view_states[update_mask] = torch.clamp(self.view_states[update_mask] + seen.view(-1, self.n_view_state_cameras),
                                                    max=1.0)

    def update_proxy_out_of_field(self, fov_proxy_mask):
        self.out_of_field[fov_proxy_mask] = 0.

    def update_proxy_supervision_occ(self, proxy_mask, signed_distances, tol=0.):
        """Pseudo ground-truth occupancy from the depth maps: a point stays 'occupied' while the fraction of the frames
        that saw it behind the surface is >= score_threshold (reference :2887-2911)."""
        self.proxy_n_inside_fov[proxy_mask] += 1
        self.proxy_n_behind_depth[proxy_mask] += (signed_distances.view(-1, 1) >= -tol).float()
        self.proxy_supervision_occ[proxy_mask] = ((self.proxy_n_behind_depth[proxy_mask] / self.proxy_n_inside_fov[proxy_mask])
                                                  >= self.score_threshold).float()

    def reset_proxy_supervision_occ(self):
        self.proxy_supervision_occ = torch.ones_like(self.proxy_supervision_occ)
