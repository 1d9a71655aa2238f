"""Drop-in `model` module: same constructor, forward/predict/loadParameters and state-dict keys as the reference
(scripts/model.py:334-432 NeuralSpeakerModel, :205-269 ResNet, :35-64 BasicBlock, :435-457 StatsPooling,
:459-501 AAMLayer), but the torch.nn modules here are only PARAMETER HOLDERS: all arithmetic runs in the sm_100a
kernels of libsvk through svk.engine.SpeakerNetEngine.  There is no torch/CPU compute fallback.

Extra keyword arguments (all default to the reference behaviour):
    precision = 'bf16' | 'fp32'   activation storage; 'fp32' is the validation mode (CUDA-core convolutions)
    impl      = None | 'tcgen05' | 'simt'   convolution implementation (default: tcgen05 for bf16, simt for fp32)
    widths, layers                channel widths / blocks per stage (the 2x-wide variant of BASELINE.json config 4)
"""
import math
import os
import sys

import torch
import torch.nn as nn

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from svk.engine import SpeakerNetEngine, run_train  # noqa: E402


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    """Parameter holder for conv-bn-relu-conv-bn (+ 1x1/s2 downsample) with a residual add (model.py:35-64)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(BasicBlock, self).__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    """Stem conv3x3(1->w0) + 4 stages of BasicBlocks, stride 2 at the start of stages 2-4 (model.py:205-244)."""

    def __init__(self, block, layers, widths=(32, 64, 128, 256)):
        super(ResNet, self).__init__()
        self.inplanes = widths[0]
        self.conv1 = nn.Conv2d(1, widths[0], kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(widths[0])
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(block, widths[0], layers[0])
        self.layer2 = self._make_layer(block, widths[1], layers[1], stride=2)
        self.layer3 = self._make_layer(block, widths[2], layers[2], stride=2)
        self.layer4 = self._make_layer(block, widths[3], layers[3], stride=2)
        for m in self.modules():                                   # model.py:222-227
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion),
            )
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)


def resnet34(widths=(32, 64, 128, 256), layers=(3, 4, 6, 3)):
    return ResNet(BasicBlock, list(layers), widths)


class StatsPooling(nn.Module):
    """'mean' or 'mean+std' temporal pooling; 'mean+std' reproduces the reference's swapped var_mean unpack, i.e.
    [unbiased variance || sqrt(mean)] (model.py:450-454)."""

    def __init__(self, pooling='mean'):
        super(StatsPooling, self).__init__()
        if pooling not in ('mean', 'mean+std'):
            raise NotImplementedError
        self.pooling = pooling


class AAMLayer(nn.Module):
    """Additive angular margin head (model.py:459-501): weight (n_classes, in_feats), xavier-normal."""

    def __init__(self, in_feats, n_classes=10, m=0.3, s=15, easy_margin=False):
        super(AAMLayer, self).__init__()
        self.m = m
        self.s = s
        self.in_feats = in_feats
        self.weight = torch.nn.Parameter(torch.FloatTensor(n_classes, in_feats), requires_grad=True)
        nn.init.xavier_normal_(self.weight, gain=1)
        if easy_margin:
            raise NotImplementedError("easy_margin is never enabled by the reference recipes")
        self.easy_margin = easy_margin
        self.cos_m = math.cos(m)
        self.sin_m = math.sin(m)
        self.th = math.cos(math.pi - m)
        self.mm = math.sin(math.pi - m) * m
        print('Initialised AAM m=%.3f s=%.3f' % (self.m, self.s))


class NeuralSpeakerModel(nn.Module):
    def __init__(self, spk_num, feat_dim=40, pooling='mean', loss='softmax', m=0.2, s=30,
                 precision='bf16', impl=None, widths=(32, 64, 128, 256), layers=(3, 4, 6, 3)):
        super(NeuralSpeakerModel, self).__init__()
        self.loss = loss
        self.res = resnet34(widths, layers)
        _feature_dim = (feat_dim + 7) // 8
        self.pool = StatsPooling(pooling=pooling)
        self.flat = nn.Flatten(1, -1)
        wl = widths[3]
        if pooling == 'mean':
            self.fc1 = nn.Linear(_feature_dim * wl, 256)
        if pooling == 'mean+std':
            self.fc1 = nn.Linear(_feature_dim * 2 * wl, 256)
        if self.loss == 'softmax':
            self.bn1 = nn.BatchNorm1d(256)
            self.fc1_relu = nn.ReLU(inplace=True)
            self.last = nn.Linear(256, spk_num)
        elif self.loss == 'AAM':
            self.last = AAMLayer(in_feats=256, n_classes=spk_num, m=m, s=s)
        elif self.loss == 'AAM-v1':
            self.bn1 = nn.BatchNorm1d(256)
            self.fc1_relu = nn.ReLU(inplace=True)
            self.last = AAMLayer(in_feats=256, n_classes=spk_num, m=m, s=s)
        else:
            raise NotImplementedError
        self.feat_dim = feat_dim
        object.__setattr__(self, "_engine", SpeakerNetEngine(self, precision=precision, impl=impl))

    # -- engine plumbing ------------------------------------------------------------------------------------
    @property
    def engine(self):
        return self._engine

    def train(self, mode=True):
        self._engine.invalidate()
        return super(NeuralSpeakerModel, self).train(mode)

    def load_state_dict(self, *args, **kwargs):
        self._engine.invalidate()
        out = super(NeuralSpeakerModel, self).load_state_dict(*args, **kwargs)
        self._engine.invalidate()
        return out

    def _check_input(self, x):
        if x.dim() != 3 or x.size(1) != self.feat_dim:
            raise ValueError("expected input of shape (B, %d, T), got %s" % (self.feat_dim, tuple(x.shape)))
        if not x.is_cuda:
            raise RuntimeError("NeuralSpeakerModel computes on CUDA only; got a %s tensor" % x.device)

    # -- reference API ----------------------------------------------------------------------------------------
    def forward(self, x, y=None):
        """(B, F, T) float [, (B,) int64 labels] -> (B, spk_num) logits (model.py:374-400)."""
        self._check_input(x)
        if self.training:
            return run_train(self._engine, x, y)
        with torch.no_grad():
            return self._engine.forward_eval(x, y=y, with_head=True)

    def forward_loss(self, x, y):
        """model(x, y) and CrossEntropyLoss()(logits, y) in one call (train_resnet.py:316-317): -> (mean loss, logits).
        AAM heads run the fused AAM-softmax-cross-entropy kernels (the backward starts from d loss: no (B, C) gradient is
        ever materialised) and attach the target ranks to the logits (`logits.svk_rank`, what accuracy() needs); the
        softmax head composes the two calls.  Training mode only; beyond the reference API."""
        self._check_input(x)
        if self.training and self.loss != "softmax" and self._engine.fused_head:
            from svk import ops
            loss, logits, rank = ops.speaker_net_train_loss(self._engine, x, y)
            logits.svk_rank = rank
            return loss, logits
        from svk.loss import CrossEntropyLoss
        logits = self(x, y)
        return CrossEntropyLoss()(logits, y), logits

    def predict(self, x, lengths=None):
        """(B, F, T) -> (B, 256) embeddings = fc1 output (model.py:402-409).  `lengths` (optional, beyond the
        reference API) gives per-row valid frame counts for zero-padded batches of different-length utterances."""
        self._check_input(x)
        with torch.no_grad():
            if self.training:
                return self._engine.forward_train(x, None, with_head=False)
            from svk import ops
            return ops.speaker_net_embed(self._engine, x, lengths)

    def loadParameters(self, loaded_state):
        """Shape-tolerant partial load that strips a leading 'module.' (model.py:415-432)."""
        self_state = self.state_dict()
        for name, param in loaded_state.items():
            origname = name
            if name not in self_state:
                name = name.replace("module.", "")
                if name not in self_state:
                    print("%s is not in the model." % origname)
                    continue
            if self_state[name].size() != loaded_state[origname].size():
                print("Wrong parameter length: %s, model: %s, loaded: %s" % (origname, self_state[name].size(),
                                                                             loaded_state[origname].size()))
                continue
            self_state[name].copy_(param)
        self._engine.invalidate()
