"""Oracle for SURVEY.md section 8f rank 2: the occupancy probability field of a whole MACARONS scene.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: /root/reference/macarons/utility/macarons_utils.py:1395-1540 (compute_scene_occupancy_probability_field),
restated cell by cell over the oracle's functional SconeOcc (oracle/scone_nets.py, a `state_dict`) and view-state
functions (oracle/view_state.py).  The scenes are duck-typed (the reference's `Scene` objects when
tests/golden/make_golden.py pins this file bit for bit against the reference's own function; any object with the same
attributes and look-up methods otherwise), the prediction camera is a pytorch3d-style camera (oracle/cameras.py).
"""
import torch

from . import scone_nets, view_state


def normalize_points_in_prediction_box(points, prediction_box_center, prediction_box_diag):
    """utility/scone_utils.py:788-796"""
    return (points - prediction_box_center) / prediction_box_diag


def scene_occupancy_field(params, occ_sd, surface_scene, proxy_scene, prediction_camera, use_supervision_occ_mask=True):
    """-> (X_world (N,3), view_harmonics (N,64), occ_probs (N,1)); updates proxy_scene.proxy_proba (:1443-1538).
    One SconeOcc forward per occupied cell and per pass of 20 000 queries, each consuming the global CPU generator
    (torch.randperm) in call order."""
    n_harm = params.n_harmonics
    X_world = torch.zeros(0, 3)
    view_harmonics = torch.zeros(0, n_harm)
    occ_probs = torch.zeros(0, 1)
    occ_mask = (proxy_scene.proxy_supervision_occ > 0.)[..., 0]
    all_fov_mask = (proxy_scene.out_of_field < 1.)[..., 0]
    if use_supervision_occ_mask:
        fovs_proxy_points = proxy_scene.proxy_points[occ_mask * all_fov_mask]
    else:
        fovs_proxy_points = proxy_scene.proxy_points[all_fov_mask]
    proxy_scene.proxy_proba[occ_mask * all_fov_mask] = 0.
    proxy_cells = proxy_scene.get_englobing_cells(fovs_proxy_points)
    base, h_polar, _ = view_state.bin_centre_harmonics(params.harmonic_degree, params.view_state_n_elev, params.view_state_n_azim)
    transform = prediction_camera.get_world_to_view_transform()
    for proxy_cell in proxy_cells:
        cell = proxy_scene.cells[proxy_scene.get_key_from_idx(proxy_cell)]
        cell_diag = torch.linalg.norm(cell.x_max - cell.x_min)
        neighbor_cells = surface_scene.get_neighboring_cells(proxy_cell)
        cell_pc_world = surface_scene.get_pt_cloud_from_cells(neighbor_cells, return_features=False)
        _, cell_X_indices = proxy_scene.get_pt_cloud_from_cells(proxy_cell, return_features=True)
        cell_X_mask = proxy_scene.get_proxy_mask_from_indices(cell_X_indices)
        if use_supervision_occ_mask:
            cell_X_mask = cell_X_mask * occ_mask
        cell_X_world = proxy_scene.proxy_points[cell_X_mask]
        box_center = transform.transform_points(cell.center.view(1, 3))
        box_diag = params.prediction_neighborhood_size * cell_diag
        if (cell_pc_world.shape[0] > 2 * 2 * params.k_for_knn) and (len(cell_X_world) > 0):
            cell_pc = normalize_points_in_prediction_box(transform.transform_points(cell_pc_world), box_center, box_diag).view(1, -1, 3)
            cell_X = normalize_points_in_prediction_box(transform.transform_points(cell_X_world), box_center, box_diag).view(1, -1, 3)
            states = view_state.move_view_state_to_view_space(
                proxy_scene.view_states[cell_X_mask.view(-1).bool()].view(1, cell_X.shape[1], params.n_view_state_cameras),
                prediction_camera, params.view_state_n_elev, params.view_state_n_azim)
            cell_vh = view_state.view_harmonics(states, base, h_polar, params.view_state_n_elev, params.view_state_n_azim)
            cell_probs = scone_nets.compute_occupancy_probability(occ_sd, cell_pc, cell_X, cell_vh,
                                                                  max_points_per_pass=20000).view(-1, 1)
            X_world = torch.vstack((X_world, cell_X_world.view(-1, 3)))
            view_harmonics = torch.vstack((view_harmonics, cell_vh.view(-1, n_harm)))
            occ_probs = torch.vstack((occ_probs, cell_probs))
            proxy_scene.proxy_proba[cell_X_mask.view(-1).bool()] = cell_probs
    oof_mask = (proxy_scene.out_of_field > 0.)[..., 0]
    oof_X_world = proxy_scene.proxy_points[oof_mask]
    X_world = torch.vstack((X_world, oof_X_world))
    view_harmonics = torch.vstack((view_harmonics, torch.zeros(len(oof_X_world), n_harm)))
    occ_probs = torch.vstack((occ_probs, proxy_scene.proxy_proba[oof_mask]))
    return X_world, view_harmonics, occ_probs
