"""network.DenseDeepGCN (network.py:256-315): DeepGCN over the 7500-vertex proxy mesh -> v_feature [1, out_channels_gcn].
Same constructor contract (an ``opt`` namespace), module tree and state-dict keys as the reference; the EdgeConv blocks run on
librnr_b200 (dropin/gcn_lib/dense/torch_vertex.py, csrc/gcn.cu)."""
import torch
from torch import nn
from torch.nn import Sequential as Seq
from torch.nn import utils

from .gcn_lib.dense import BasicConv, DenseDilatedKnnGraph, DenseDynBlock4D, GraphConv4D, ResDynBlock4D


class DenseDeepGCN(nn.Module):
    def __init__(self, opt):
        super().__init__()
        channels, k = opt.n_filters, opt.kernel_size
        act, norm, bias = opt.act_type, opt.norm_type, opt.bias
        eps, stochastic, conv = opt.epsilon, opt.stochastic, opt.conv_type
        growth = channels
        self.n_blocks = opt.n_blocks
        self.knn = DenseDilatedKnnGraph(k, 1, stochastic, eps)
        self.head = GraphConv4D(opt.in_channels, channels, conv, act, norm, bias)
        kind = opt.block_type.lower()
        if kind == 'res':
            self.backbone = Seq(*[ResDynBlock4D(channels, k, 1 + i, conv, act, norm, bias, stochastic, eps)
                                  for i in range(self.n_blocks - 1)])
        elif kind == 'dense':
            self.backbone = Seq(*[DenseDynBlock4D(channels + growth * i, growth, k, 1 + i, conv, act, norm, bias, stochastic, eps)
                                  for i in range(self.n_blocks - 1)])
        else:
            raise NotImplementedError('{} is not implemented. Please check.\n'.format(opt.block_type))
        wide = channels + growth * (self.n_blocks - 1)
        self.fusion_block = BasicConv([wide, 1024], act, None, bias)
        # constructed (and present in checkpoints) but never called by the reference either (network.py:285-286)
        self.prediction = Seq(BasicConv([1 + wide, 512, 256], act, None, bias), BasicConv([256, 64], act, None, bias))
        self.linear = Seq(utils.spectral_norm(nn.Linear(opt.num_v_gcn, 2048)),
                          utils.spectral_norm(nn.Linear(2048, opt.out_channels_gcn)))
        self.model_init()

    def model_init(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                m.weight.requires_grad = True
                if m.bias is not None:
                    m.bias.data.zero_()
                    m.bias.requires_grad = True
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.normal_(1.0, 0.02)
                m.bias.data.fill_(0)

    def forward(self, inputs):
        data = torch.cat((inputs.pos, inputs.x), 1).unsqueeze(0).unsqueeze(-1)           # [1, V, 6, 1]
        feats = [self.head(data.transpose(2, 1), self.knn(data[:, :, 0:3]))]
        for block in self.backbone:
            feats.append(block(feats[-1]))
        feats = torch.cat(feats, 1)
        fusion = torch.max(self.fusion_block(feats), 1, keepdim=True)[0]                 # max over channels -> [1,1,V,1]
        return self.linear(fusion.view(-1)).unsqueeze(0)
