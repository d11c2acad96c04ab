"""ManyDepth: multi-frame depth prediction (reference macarons/networks/ManyDepth.py:33-790).

Same class names, constructor arguments, attribute tree / state_dict keys and forward signature as the reference.
The modules hold the parameters; `ManyDepth.forward` composes the source cameras from the ground-truth relative poses
(ManyDepth.py:740-750) and then runs ONE fused CUDA forward through the C ABI (csrc/depth.cu: im2col gathers + tcgen05
linear layers for every convolution, a single plane-sweep cost-volume kernel).  Inference only: BatchNorm layers are
folded with their running statistics, `learn_pose=True` (PoseDecoder) is outside the NBV path.  CUDA tensors only.
"""
import torch
from torch import nn

from .. import netpack, ops
from ..utility import rotations
from .Attention import _FusedOnly

input_height, input_width, input_channels = 256, 456, 3       # reference :19-21
d_min, d_max, n_alpha, n_depth = 0.5, 750, 2, 96               # reference :23-26
pose_factor, learn_pose = 100., False                          # reference :28-29


class BasicBlock(_FusedOnly):
    """torchvision.models.resnet.BasicBlock parameter layout (conv1, bn1, conv2, bn2[, downsample.{0,1}])."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        self.stride = stride


class ResNet18Trunk(nn.Module):
    """The part of torchvision's resnet18 the reference uses (conv1, bn1, relu, maxpool, layer1-4), same attribute
    names, built locally (the reference downloads it through torch.hub, ManyDepth.py:764-771)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = nn.Sequential(BasicBlock(64, 64), BasicBlock(64, 64))
        self.layer2 = nn.Sequential(BasicBlock(64, 128, 2), BasicBlock(128, 128))
        self.layer3 = nn.Sequential(BasicBlock(128, 256, 2), BasicBlock(256, 256))
        self.layer4 = nn.Sequential(BasicBlock(256, 512, 2), BasicBlock(512, 512))


class FeatureExtractor(_FusedOnly):
    def __init__(self, resnet_model):
        super().__init__()
        self.conv1, self.bn1, self.relu = resnet_model.conv1, resnet_model.bn1, resnet_model.relu
        self.maxpool, self.layer = resnet_model.maxpool, resnet_model.layer1


class CostVolumeBuilder(_FusedOnly):
    def __init__(self, height, width, feature_height, feature_width, feature_channels, n_alpha, d_min, d_max, n_depth,
                 output_channels, kernel_size=3, stride=1, padding=1):
        super().__init__()
        self.height, self.width = height, width
        self.feature_height, self.feature_width, self.feature_channels = feature_height, feature_width, feature_channels
        self.n_alpha, self.d_min, self.d_max, self.n_depth = n_alpha, d_min, d_max, n_depth
        self.conv_reduce = nn.Conv2d(feature_channels + n_depth, output_channels, kernel_size, stride, padding)
        self.relu = nn.ReLU()


class ExpansionLayer(_FusedOnly):
    def __init__(self, input_channels, inner_channels, output_channels, output_size, additional_channels=None,
                 kernel_size=3, stride=1, padding=1):
        super().__init__()
        self.upconv = nn.ConvTranspose2d(input_channels, inner_channels, kernel_size, stride, padding)
        self.upelu = nn.ELU()
        total = inner_channels + (additional_channels or 0)
        self.iconv = nn.Conv2d(total, output_channels, kernel_size, stride, padding, padding_mode='reflect')
        self.ielu = nn.ELU()
        self.input_channels, self.output_channels, self.output_size = input_channels, output_channels, output_size
        self.inner_channels, self.additional_channels = inner_channels, additional_channels
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding


class DisparityLayer(_FusedOnly):
    def __init__(self, input_channels):
        super().__init__()
        self.channels = input_channels
        self.conv = nn.Conv2d(input_channels, 1, 3, 1, 1, padding_mode='reflect')
        self.sigmoid = nn.Sigmoid()


class DepthDecoder(_FusedOnly):
    def __init__(self, feature_extractor, resnet_model, input_height=input_height, input_width=input_width,
                 input_channels=input_channels, n_alpha=n_alpha, d_min=d_min, d_max=d_max, n_depth=n_depth,
                 use_input_image_in_skip_connection=True):
        super().__init__()
        if not use_input_image_in_skip_connection or input_channels != 3:
            raise NotImplementedError("the fused depth forward uses the RGB frame in the last skip connection")
        self.height, self.width, self.channels = input_height, input_width, input_channels
        self.use_input_image_in_skip_connection = use_input_image_in_skip_connection
        self.feature_extractor = feature_extractor
        self.cost_volume_builder = CostVolumeBuilder(input_height, input_width, input_height // 4, input_width // 4, 64,
                                                     n_alpha, d_min, d_max, n_depth, 64)
        self.resnet_layer_2, self.resnet_layer_3, self.resnet_layer_4 = resnet_model.layer2, resnet_model.layer3, resnet_model.layer4
        size = lambda d: (input_height // d, input_width // d + (input_width % d > 0))
        self.expansion5 = ExpansionLayer(512, 256, 256, size(16), additional_channels=256)
        self.expansion4 = ExpansionLayer(256, 128, 128, size(8), additional_channels=128)
        self.disp4 = DisparityLayer(128)
        self.expansion3 = ExpansionLayer(128, 64, 64, size(4), additional_channels=64)
        self.disp3 = DisparityLayer(64)
        self.expansion2 = ExpansionLayer(64, 32, 32, size(2), additional_channels=64)
        self.disp2 = DisparityLayer(32)
        self.expansion1 = ExpansionLayer(32, 16, 16, (input_height, input_width), additional_channels=3)
        self.disp1 = DisparityLayer(16)


class ManyDepth(nn.Module):
    def __init__(self, depth_decoder, pose_decoder, pose_factor=pose_factor, learn_pose=learn_pose):
        super().__init__()
        if learn_pose:
            raise NotImplementedError("learn_pose=True (PoseDecoder) is not on the NBV scoring path (reference default False)")
        self.depth_decoder = depth_decoder
        self.pose_factor, self.learn_pose = pose_factor, learn_pose
        self.input_height, self.input_width = depth_decoder.height, depth_decoder.width
        self.d_min = depth_decoder.cost_volume_builder.d_min
        self.d_max = depth_decoder.cost_volume_builder.d_max
        self.n_depth = depth_decoder.cost_volume_builder.n_depth

    def forward(self, x, x_alpha, R, T, zfar, device, gt_pose=None):
        """x (B,3,H,W), x_alpha (B,n_alpha,3,H,W), R (B,3,3), T (B,3), zfar (B,), gt_pose (B,n_alpha,6)
        -> (pose, disp1 (B,1,H,W), disp2, disp3, disp4)   [reference ManyDepth.py:719-758]"""
        if self.d_max != zfar[0].item():      # (a host synchronisation per frame, as in the reference :728)
            raise NameError("Model variable d_max is different from the provided zfar.\n"
                            "Please check that d_min and d_max are respectively equal to parameters znear and zfar.")
        if gt_pose is None:
            raise NameError("Input gt_pose is missing!The parameter 'learn_pose' is set to False. "
                            "Consequently, Model must take ground truth poses as an input.")
        if self.training:
            raise NotImplementedError("the fused depth forward folds BatchNorm with its running statistics: call .eval()")
        ops.refuse_grad("ManyDepth.forward", x, x_alpha, module=self)
        B, n_a = x.shape[0], x_alpha.shape[1]
        pose = gt_pose
        R_alpha, T_alpha = rotations.relative_cameras(R, T, pose, self.pose_factor)
        cam = torch.cat((torch.cat((R.reshape(B, 1, 9), T.reshape(B, 1, 3), zfar.reshape(B, 1, 1).to(R.dtype)), dim=-1),
                         torch.cat((R_alpha.reshape(B, n_a, 9), T_alpha.reshape(B, n_a, 3),
                                    zfar.reshape(B, 1, 1).expand(-1, n_a, -1).to(R.dtype)), dim=-1)), dim=1)
        disps = ops.manydepth_forward(netpack.pack_manydepth(self), x, x_alpha, cam.to(torch.float32).contiguous())
        return (pose,) + disps


def create_many_depth_model(device, learn_pose=learn_pose, pretrained_resnet_path=None, save_resnet=False):
    """reference :761-805 without the torch.hub download: a locally built ResNet-18 trunk (optionally loaded from a
    torchvision resnet18 state dict at `pretrained_resnet_path`)."""
    resnet = ResNet18Trunk()
    if pretrained_resnet_path is not None:
        resnet.load_state_dict(torch.load(pretrained_resnet_path, map_location="cpu"), strict=False)
    fe = FeatureExtractor(resnet)
    dd = DepthDecoder(fe, resnet, input_height=input_height, input_width=input_width, input_channels=input_channels)
    return ManyDepth(depth_decoder=dd, pose_decoder=None, learn_pose=learn_pose).to(device).eval()
