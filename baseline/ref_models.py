"""Plain-PyTorch definitions of the two benchmark architectures, used ONLY by the reference arm
(`bench.py --impl reference`) and as numerics oracles in tests.  Self-contained: imports nothing
from coinstac_dinunet_b200.  The architectures are documented in DESIGN.md and mirrored by
coinstac_dinunet_b200.models.{fsnet,vbmnet} (same layer order => same parameter order/shapes).
"""
from torch import nn


class RefFSNet(nn.Module):
    def __init__(self, in_size=66, hidden_sizes=(256, 128, 64, 32), out_size=2):
        super().__init__()
        layers, prev = [], in_size
        for h in hidden_sizes:
            layers += [nn.Linear(prev, h), nn.BatchNorm1d(h), nn.ReLU(inplace=True)]
            prev = h
        self.features = nn.Sequential(*layers)
        self.classifier = nn.Linear(prev, out_size)

    def forward(self, x):
        return self.classifier(self.features(x.reshape(x.shape[0], -1)))


class _RefConvBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, kernel_size=3, padding=1, bias=False)
        self.bn = nn.BatchNorm3d(cout)
        self.act = nn.ReLU(inplace=True)
        self.pool = nn.MaxPool3d(2)

    def forward(self, x):
        return self.pool(self.act(self.bn(self.conv(x))))


class RefVBMNet(nn.Module):
    def __init__(self, in_ch=1, num_class=2, channels=(16, 32, 64, 128, 256), head=(256, 64),
                 input_shape=(121, 145, 121)):
        super().__init__()
        chans = [in_ch, *channels]
        self.blocks = nn.Sequential(*[_RefConvBlock(a, b) for a, b in zip(chans[:-1], chans[1:])])
        d, h, w = input_shape
        for _ in channels:
            d, h, w = d // 2, h // 2, w // 2
        dims = [channels[-1] * d * h * w, *head]
        fc = []
        for a, b in zip(dims[:-1], dims[1:]):
            fc += [nn.Linear(a, b), nn.ReLU(inplace=True)]
        self.head = nn.Sequential(*fc)
        self.classifier = nn.Linear(dims[-1], num_class)

    def forward(self, x):
        if x.dim() == 4:
            x = x.unsqueeze(1)
        return self.classifier(self.head(self.blocks(x).flatten(1)))
