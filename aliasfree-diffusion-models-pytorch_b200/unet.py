"""UNet wiring of the reference (modules/ddpm_models.py:41-298), variants 0-3.

The reference UNet is Python glue around the blocks; it is restated here only because the
GPU box has no copy of the reference to import.  Attribute names (inc, down1, sa1, ...,
outc, label_emb) and therefore state_dict keys are identical; channel widths follow the
reference's ``image_size * {1, 2, 4, 8}`` rule.  Variant 4 (GroupNorm at 2x resolution, an
upstream experiment) is supported through the standalone resampling kernels (no fusion).
"""
import torch
import torch.nn as nn

from . import blocks as B

# variant -> (DoubleConv class, Down class, Up class, needs f_settings)
_VARIANTS = {
    0: (B.DoubleConv, B.Down, B.Up, False),
    1: (B.DoubleConv, B.Down_FF, B.Up_FF, True),       # Config B
    2: (B.DoubleConv_F, B.Down_F, B.Up_F, True),       # Config C
    3: (B.DoubleConv_F, B.Down_FFF, B.Up_FFF, True),   # Config D
    4: (B.DoubleConv_F4, B.Down_F4, B.Up_F4, True),    # experimental upstream: GroupNorm on the 2x grid
}


class UNet(nn.Module):
    def __init__(self, c_in=3, c_out=3, image_size=64, time_dim=256, device="cuda", f_settings=None,
                 num_classes=None, variant=0):
        super().__init__()
        if variant not in _VARIANTS:
            raise ValueError("variant value must be between 0 and 4")
        conv_cls, down_cls, up_cls, filtered = _VARIANTS[variant]
        if filtered and f_settings is None:
            raise ValueError("f_settings is empty")
        self.device, self.time_dim, self.image_size, self.f_settings = device, time_dim, image_size, f_settings
        self.variant = variant
        s = int(image_size)
        stage_kw = {"f_settings": f_settings} if filtered else {}
        conv_kw = {"f_settings": f_settings} if issubclass(conv_cls, B.DoubleConv_F) else {}

        self.inc = conv_cls(c_in, s, **conv_kw)
        self.down1 = down_cls(s, 2 * s, **stage_kw)
        self.sa1 = B.SelfAttention(2 * s, s // 2)
        self.down2 = down_cls(2 * s, 4 * s, **stage_kw)
        self.sa2 = B.SelfAttention(4 * s, s // 4)
        self.down3 = down_cls(4 * s, 4 * s, **stage_kw)
        self.sa3 = B.SelfAttention(4 * s, s // 8)
        self.bot1 = conv_cls(4 * s, 8 * s, **conv_kw)
        self.bot2 = conv_cls(8 * s, 8 * s, **conv_kw)
        self.bot3 = conv_cls(8 * s, 4 * s, **conv_kw)
        self.up1 = up_cls(8 * s, 2 * s, **stage_kw)
        self.sa4 = B.SelfAttention(2 * s, s // 4)
        self.up2 = up_cls(4 * s, s, **stage_kw)
        self.sa5 = B.SelfAttention(s, s // 2)
        self.up3 = up_cls(2 * s, s, **stage_kw)
        self.sa6 = B.SelfAttention(s, s)
        self.outc = nn.Conv2d(s, c_out, kernel_size=1)
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, time_dim)

    def pos_encoding(self, t, channels):
        """Sinusoidal time encoding, modules/ddpm_models.py:261-269. ``t`` is [B, 1] float."""
        expo = torch.arange(0, channels, 2, device=t.device).float() / channels
        ang = t * (1.0 / (10000 ** expo))
        return torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1)

    def forward(self, x, t, y=None):
        t = self.pos_encoding(t.unsqueeze(-1).float(), self.time_dim)
        if y is not None:
            t = t + self.label_emb(y)
        x1 = self.inc(x)
        x2 = self.sa1(self.down1(x1, t))
        x3 = self.sa2(self.down2(x2, t))
        x4 = self.sa3(self.down3(x3, t))
        x4 = self.bot3(self.bot2(self.bot1(x4)))
        h = self.sa4(self.up1(x4, x3, t))
        h = self.sa5(self.up2(h, x2, t))
        h = self.sa6(self.up3(h, x1, t))
        return self.outc(h)
