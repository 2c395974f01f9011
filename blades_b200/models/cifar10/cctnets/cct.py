"""CCT (Compact Convolutional Transformer) family -- reference cctnets/cct.py:33-352.
All 22 ``cct_*`` factories are generated from one spec table."""
import torch.nn as nn

from .core import Tokenizer, TransformerClassifier, register_model

__all__ = ["CCT", "cct_2", "cct_4", "cct_6", "cct_7", "cct_14"]


class CCT(nn.Module):
    def __init__(self, img_size=224, embedding_dim=768, n_input_channels=3, n_conv_layers=1, kernel_size=7,
                 stride=2, padding=3, pooling_kernel_size=3, pooling_stride=2, pooling_padding=1, dropout=0.,
                 attention_dropout=0.1, stochastic_depth=0.1, num_layers=14, num_heads=6, mlp_ratio=4.0,
                 num_classes=1000, positional_embedding='learnable', *args, **kwargs):
        super().__init__()
        self.tokenizer = Tokenizer(n_input_channels=n_input_channels, n_output_channels=embedding_dim,
                                   kernel_size=kernel_size, stride=stride, padding=padding,
                                   pooling_kernel_size=pooling_kernel_size, pooling_stride=pooling_stride,
                                   pooling_padding=pooling_padding, max_pool=True, activation=nn.ReLU,
                                   n_conv_layers=n_conv_layers, conv_bias=False)
        self.classifier = TransformerClassifier(
            sequence_length=self.tokenizer.sequence_length(n_channels=n_input_channels, height=img_size,
                                                           width=img_size),
            embedding_dim=embedding_dim, seq_pool=True, dropout=dropout, attention_dropout=attention_dropout,
            stochastic_depth=stochastic_depth, num_layers=num_layers, num_heads=num_heads,
            mlp_ratio=mlp_ratio, num_classes=num_classes, positional_embedding=positional_embedding)

    def forward(self, x):
        return self.classifier(self.tokenizer(x))


#: depth -> (num_layers, num_heads, mlp_ratio, embedding_dim)   (reference cct.py:121-143)
_SIZES = {2: (2, 2, 1, 128), 4: (4, 2, 1, 128), 6: (6, 4, 2, 256), 7: (7, 4, 2, 256), 14: (14, 6, 3, 384)}


def _cct(arch, pretrained, progress, num_layers, num_heads, mlp_ratio, embedding_dim, kernel_size=3,
         stride=None, padding=None, positional_embedding='learnable', *args, **kwargs):
    if pretrained:
        raise RuntimeError(f'Variant {arch}: pretrained weights need network access (not available).')
    stride = stride if stride is not None else max(1, (kernel_size // 2) - 1)
    padding = padding if padding is not None else max(1, (kernel_size // 2))
    return CCT(num_layers=num_layers, num_heads=num_heads, mlp_ratio=mlp_ratio, embedding_dim=embedding_dim,
               kernel_size=kernel_size, stride=stride, padding=padding,
               positional_embedding=positional_embedding, *args, **kwargs)


def _size_factory(depth):
    L, H, R, E = _SIZES[depth]

    def f(arch, pretrained, progress, *args, **kwargs):
        return _cct(arch, pretrained, progress, num_layers=L, num_heads=H, mlp_ratio=R, embedding_dim=E,
                    *args, **kwargs)
    f.__name__ = f"cct_{depth}"
    return f


cct_2, cct_4, cct_6, cct_7, cct_14 = (_size_factory(d) for d in (2, 4, 6, 7, 14))

# name -> (depth, kernel, n_conv, img_size, pos-emb, classes)      (reference cct.py:146-352)
_VARIANTS = {
    'cct_2_3x2_32': (2, 3, 2, 32, 'learnable', 10), 'cct_2_3x2_32_sine': (2, 3, 2, 32, 'sine', 10),
    'cct_4_3x2_32': (4, 3, 2, 32, 'learnable', 10), 'cct_4_3x2_32_sine': (4, 3, 2, 32, 'sine', 10),
    'cct_6_3x1_32': (6, 3, 1, 32, 'learnable', 10), 'cct_6_3x1_32_sine': (6, 3, 1, 32, 'sine', 10),
    'cct_6_3x2_32': (6, 3, 2, 32, 'learnable', 10), 'cct_6_3x2_32_sine': (6, 3, 2, 32, 'sine', 10),
    'cct_7_3x1_32': (7, 3, 1, 32, 'learnable', 10), 'cct_7_3x1_32_sine': (7, 3, 1, 32, 'sine', 10),
    'cct_7_3x1_32_c100': (7, 3, 1, 32, 'learnable', 100), 'cct_7_3x1_32_sine_c100': (7, 3, 1, 32, 'sine', 100),
    'cct_7_3x2_32': (7, 3, 2, 32, 'learnable', 10), 'cct_7_3x2_32_sine': (7, 3, 2, 32, 'sine', 10),
    'cct_7_7x2_224': (7, 7, 2, 224, 'learnable', 102), 'cct_7_7x2_224_sine': (7, 7, 2, 224, 'sine', 102),
    'cct_14_7x2_224': (14, 7, 2, 224, 'learnable', 1000), 'cct_14_7x2_384': (14, 7, 2, 384, 'learnable', 1000),
    'cct_14_7x2_384_fl': (14, 7, 2, 384, 'learnable', 102),
}


def _variant(name, depth, ks, nconv, img, pe, ncls):
    base = {2: cct_2, 4: cct_4, 6: cct_6, 7: cct_7, 14: cct_14}[depth]

    def f(pretrained=False, progress=False, img_size=img, positional_embedding=pe, num_classes=ncls,
          *args, **kwargs):
        return base(name, pretrained, progress, kernel_size=ks, n_conv_layers=nconv, img_size=img_size,
                    positional_embedding=positional_embedding, num_classes=num_classes, *args, **kwargs)
    f.__name__ = name
    f.__doc__ = f"CCT-{depth}/{ks}x{nconv}, {img}px, {pe} positional embedding, {ncls} classes."
    return register_model(f)


for _n, _spec in _VARIANTS.items():
    globals()[_n] = _variant(_n, *_spec)
    __all__.append(_n)
from .utils.helpers import fc_check, pe_check  # noqa: F401,E402  (import-path parity with the reference module)
