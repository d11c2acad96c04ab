"""Oracle for row a14 of SURVEY.md section 8: the multi-frame depth network (ManyDepth forward, inference mode:
BatchNorm in eval mode, ground-truth relative poses).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional restatement over a plain `state_dict` with the same torch
CPU kernels in the same order as the reference modules; the pytorch3d camera calls go through oracle/cameras.py
(restated, parity unpinned there).  `tests/golden/make_golden.py` asserts bit-equality with the reference's own
ManyDepth.forward (run with the same camera stand-ins) before writing the fixtures.

Reference (paths relative to /root/reference/macarons/networks/ManyDepth.py):
  :33-50   FeatureExtractor.forward          :111-144 CostVolumeBuilder.reproject_depth_map
  :146-205 CostVolumeBuilder.warp            :207-305 CostVolumeBuilder.forward
  :349-365 ExpansionLayer.forward            :383-384 DisparityLayer.forward
  :474-531 DepthDecoder.forward              :719-758 ManyDepth.forward
torchvision resnet18 BasicBlock (conv3x3-bn-relu-conv3x3-bn (+downsample) -add-relu).
"""
import torch
import torch.nn.functional as F

from . import cameras as cams

N_DEPTH, D_MIN, D_MAX = 96, 0.5, 750.0       # ManyDepth.py:24-26


def _bn(sd, name, x):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.1, 1e-5)


def _basic_block(sd, name, x, stride):
    out = F.relu(_bn(sd, name + ".bn1", F.conv2d(x, sd[name + ".conv1.weight"], None, stride, 1)))
    out = _bn(sd, name + ".bn2", F.conv2d(out, sd[name + ".conv2.weight"], None, 1, 1))
    if (name + ".downsample.0.weight") in sd:
        x = _bn(sd, name + ".downsample.1", F.conv2d(x, sd[name + ".downsample.0.weight"], None, stride, 0))
    return F.relu(out + x)


def _layer(sd, name, x, stride):
    return _basic_block(sd, name + ".1", _basic_block(sd, name + ".0", x, stride), 1)


def feature_extractor(sd, x, p="depth_decoder.feature_extractor"):
    """-> (conv1 after bn+relu, layer1 output)   [ManyDepth.py:486-491]"""
    conv1 = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, 2, 3)))
    return conv1, _layer(sd, p + ".layer", F.max_pool2d(conv1, 3, 2, 1), 1)


def _transpose_channels(img, channel_is_at_the_end):
    if channel_is_at_the_end:
        return torch.transpose(0. + torch.transpose(img, -1, -2), -2, -3)
    return torch.transpose(0. + torch.transpose(img, -3, -2), -2, -1)


def cost_volume(x, R, T, zfar, x_alpha, R_alpha, T_alpha, zfar_alpha, height, width):
    """ManyDepth.py:207-297 -> (B, n_depth, fh, fw): mean over source frames of the warped source features,
    L1 distance to the target features over channels / n_channels."""
    B, n_alpha = x.shape[0], x_alpha.shape[1]
    C, fh, fw = x.shape[1], x.shape[2], x.shape[3]
    x_tab = torch.Tensor([[i for _ in range(width)] for i in range(height)])
    y_tab = torch.Tensor([[j for j in range(width)] for _ in range(height)])
    depth_bins = torch.linspace(D_MIN, D_MAX, N_DEPTH)
    cam = cams.FoVPerspectiveCameras(R=R.view(B, 1, 3, 3).expand(-1, N_DEPTH, -1, -1).contiguous().view(-1, 3, 3),
                                     T=T.view(B, 1, 3).expand(-1, N_DEPTH, -1).contiguous().view(-1, 3),
                                     zfar=zfar.view(B, 1).expand(-1, N_DEPTH).contiguous().view(-1))
    cam_a = cams.FoVPerspectiveCameras(
        R=R_alpha.view(B, 1, n_alpha, 3, 3).expand(-1, N_DEPTH, -1, -1, -1).contiguous().view(-1, 3, 3),
        T=T_alpha.view(B, 1, n_alpha, 3).expand(-1, N_DEPTH, -1, -1).contiguous().view(-1, 3),
        zfar=zfar_alpha.view(B, 1, n_alpha).expand(-1, N_DEPTH, -1).contiguous().view(-1))
    depth = depth_bins.view(1, -1, 1, 1, 1).expand(B, -1, height, width, -1).contiguous().view(-1, height, width, 1)
    n_img = depth.shape[0]
    m = min(width, height)
    ndc_x = width / m - (y_tab / (m - 1)) * 2
    ndc_y = height / m - (x_tab / (m - 1)) * 2
    ndc = torch.cat((ndc_x.view(1, -1, 1).expand(n_img, -1, -1), ndc_y.view(1, -1, 1).expand(n_img, -1, -1),
                     depth.view(n_img, -1, 1)), dim=-1).view(n_img, height * width, 3)
    world = cam.unproject_points(ndc, scaled_depth_input=False)
    target = world.view(B, N_DEPTH, 1, height, width, 3).expand(-1, -1, n_alpha, -1, -1, -1).contiguous().view(-1, height, width, 3)
    src = x_alpha.view(B, 1, n_alpha, C, fh, fw).expand(-1, N_DEPTH, -1, -1, -1, -1).contiguous().view(-1, C, fh, fw)
    n_w = src.shape[0]
    screen = cam_a.get_full_projection_transform().transform_points(target.view(n_w, -1, 3), eps=1e-8)
    factor = -1 * min(fw, fh)
    screen[..., 0] = factor / fw * screen[..., 0]
    screen[..., 1] = factor / fh * screen[..., 1]
    screen = screen[..., :2].view(n_w, height, width, 2)
    screen = _transpose_channels(screen, True)
    screen = F.interpolate(screen, size=(fh, fw), mode='bicubic')
    screen = _transpose_channels(screen, False)
    warped = F.grid_sample(input=src, grid=screen, mode='bilinear', padding_mode='zeros', align_corners=False)
    warped = torch.mean(warped.view(B, N_DEPTH, n_alpha, C, fh, fw), dim=-4)
    return torch.linalg.norm(warped - x.view(B, 1, C, fh, fw).expand(-1, N_DEPTH, -1, -1, -1), dim=2, ord=1) / C


def _expansion(sd, name, x, x_add, output_size):
    """ManyDepth.py:349-365: transposed conv + ELU, nearest up-sampling, skip concatenation, reflect-padded conv + ELU."""
    res = F.elu(F.conv_transpose2d(x, sd[name + ".upconv.weight"], sd[name + ".upconv.bias"], 1, 1))
    res = F.interpolate(input=res, size=output_size, mode='nearest')
    if x_add is not None:
        res = torch.cat((res, x_add), dim=-3)
    res = F.conv2d(F.pad(res, (1, 1, 1, 1), mode='reflect'), sd[name + ".iconv.weight"], sd[name + ".iconv.bias"])
    return F.elu(res)


def _disp(sd, name, x):
    return torch.sigmoid(F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), sd[name + ".conv.weight"], sd[name + ".conv.bias"]))


def depth_decoder(sd, x, R, T, zfar, x_alpha, R_alpha, T_alpha, zfar_alpha, return_stages=False):
    """ManyDepth.py:474-531 -> disp1..disp4."""
    p = "depth_decoder"
    B, n_alpha, H, W = x.shape[0], x_alpha.shape[1], x.shape[2], x.shape[3]
    conv1, layer1 = feature_extractor(sd, x)
    _, layer1_a = feature_extractor(sd, x_alpha.reshape(-1, 3, H, W))
    layer1_a = layer1_a.view(B, n_alpha, 64, H // 4, W // 4)
    cv = cost_volume(layer1, R, T, zfar, layer1_a, R_alpha, T_alpha, zfar_alpha, H, W)
    red = F.relu(F.conv2d(torch.cat((layer1, cv), dim=-3), sd[p + ".cost_volume_builder.conv_reduce.weight"],
                          sd[p + ".cost_volume_builder.conv_reduce.bias"], 1, 1))
    layer2 = _layer(sd, p + ".resnet_layer_2", red, 2)
    layer3 = _layer(sd, p + ".resnet_layer_3", layer2, 2)
    layer4 = _layer(sd, p + ".resnet_layer_4", layer3, 2)
    up = lambda d: (H // d, W // d + (W % d > 0))
    iconv5 = _expansion(sd, p + ".expansion5", layer4, layer3, up(16))
    iconv4 = _expansion(sd, p + ".expansion4", iconv5, layer2, up(8))
    iconv3 = _expansion(sd, p + ".expansion3", iconv4, layer1, up(4))
    iconv2 = _expansion(sd, p + ".expansion2", iconv3, conv1, up(2))
    iconv1 = _expansion(sd, p + ".expansion1", iconv2, x, (H, W))
    out = (_disp(sd, p + ".disp1", iconv1), _disp(sd, p + ".disp2", iconv2), _disp(sd, p + ".disp3", iconv3),
           _disp(sd, p + ".disp4", iconv4))
    if return_stages:
        return out, {"conv1": conv1, "layer1": layer1, "cost_volume": cv, "conv_reduce": red, "layer4": layer4, "iconv1": iconv1}
    return out


def relative_cameras(R, T, pose, pose_factor=100.):
    """ManyDepth.py:740-750: source-frame cameras from the 6-vector relative poses (B, n_alpha, 6)."""
    B, n_alpha = pose.shape[0], pose.shape[1]
    rel_R = cams.axis_angle_to_matrix(pose_factor * pose[..., 3:])
    rel_T = pose_factor * pose[..., :3]
    eR = R.view(B, 1, 3, 3).expand(-1, n_alpha, -1, -1)
    eT = T.view(B, 1, 3).expand(-1, n_alpha, -1)
    R_alpha = eR @ rel_R
    T_alpha = rel_T + cams.quaternion_apply(cams.matrix_to_quaternion(rel_R.transpose(dim0=-1, dim1=-2)), eT)
    return R_alpha, T_alpha


def many_depth_forward(sd, x, x_alpha, R, T, zfar, gt_pose, pose_factor=100., return_stages=False):
    """ManyDepth.py:719-758 with learn_pose=False -> (pose, disp1, disp2, disp3, disp4)."""
    B, n_alpha = x.shape[0], x_alpha.shape[1]
    R_alpha, T_alpha = relative_cameras(R, T, gt_pose, pose_factor)
    zfar_alpha = zfar.view(B, 1).expand(-1, n_alpha).contiguous()
    res = depth_decoder(sd, x, R, T, zfar, x_alpha, R_alpha, T_alpha, zfar_alpha, return_stages=return_stages)
    if return_stages:
        return (gt_pose,) + res[0], res[1]
    return (gt_pose,) + res


def depth_from_disparity(disp, znear=D_MIN, zfar=D_MAX):
    """utility/depth_model_utils.py:844-848"""
    a, b = 1. / znear - 1. / zfar, 1. / zfar
    return 1. / (a * disp + b)
