"""CVT (Compact Vision Transformer: patch tokenizer + seq-pool) -- reference cctnets/cvt.py:17-198."""
import torch.nn as nn

from .core import Tokenizer, TransformerClassifier, register_model

__all__ = ["CVT"]


class _PatchModel(nn.Module):
    _seq_pool = True

    def __init__(self, img_size=224, embedding_dim=768, n_input_channels=3, kernel_size=16, dropout=0.,
                 attention_dropout=0.1, stochastic_depth=0.1, num_layers=14, num_heads=6, mlp_ratio=4.0,
                 num_classes=1000, positional_embedding='learnable', *args, **kwargs):
        super().__init__()
        assert img_size % kernel_size == 0, \
            f"Image size ({img_size}) has to be divisible by patch size ({kernel_size})"
        self.tokenizer = Tokenizer(n_input_channels=n_input_channels, n_output_channels=embedding_dim,
                                   kernel_size=kernel_size, stride=kernel_size, padding=0, max_pool=False,
                                   activation=None, n_conv_layers=1, conv_bias=True)
        self.classifier = TransformerClassifier(
            sequence_length=self.tokenizer.sequence_length(n_channels=n_input_channels, height=img_size,
                                                           width=img_size),
            embedding_dim=embedding_dim, seq_pool=self._seq_pool, dropout=dropout,
            attention_dropout=attention_dropout, stochastic_depth=stochastic_depth, num_layers=num_layers,
            num_heads=num_heads, mlp_ratio=mlp_ratio, num_classes=num_classes,
            positional_embedding=positional_embedding)

    def forward(self, x):
        return self.classifier(self.tokenizer(x))


class CVT(_PatchModel):
    _seq_pool = True


_SIZES = {2: (2, 2, 1, 128), 4: (4, 2, 1, 128), 6: (6, 4, 2, 256), 7: (7, 4, 2, 256), 8: (8, 4, 2, 256)}


def _make_family(cls, prefix, depths, force_learnable=False):
    out = {}

    def _build(arch, pretrained, progress, num_layers, num_heads, mlp_ratio, embedding_dim, kernel_size=4,
               positional_embedding='learnable', *args, **kwargs):
        if force_learnable:
            positional_embedding = 'learnable'
        return cls(num_layers=num_layers, num_heads=num_heads, mlp_ratio=mlp_ratio,
                   embedding_dim=embedding_dim, kernel_size=kernel_size,
                   positional_embedding=positional_embedding, *args, **kwargs)
    out[f"_{prefix}"] = _build
    for d in depths:
        L, H, R, E = _SIZES[d]

        def size_f(*args, _L=L, _H=H, _R=R, _E=E, **kwargs):
            return _build(*args, num_layers=_L, num_heads=_H, mlp_ratio=_R, embedding_dim=_E, **kwargs)
        size_f.__name__ = f"{prefix}_{d}"
        out[size_f.__name__] = size_f
    for d in (2, 4, 6, 7):
        for pe in ('learnable', 'sine'):
            name = f"{prefix}_{d}_4_32" + ("_sine" if pe == 'sine' else "")

            def var_f(pretrained=False, progress=False, img_size=32, positional_embedding=pe, num_classes=10,
                      *args, _name=name, _d=d, **kwargs):
                return out[f"{prefix}_{_d}"](_name, pretrained, progress, kernel_size=4, img_size=img_size,
                                             positional_embedding=positional_embedding,
                                             num_classes=num_classes, *args, **kwargs)
            var_f.__name__ = name
            out[name] = register_model(var_f)
    return out


_fam = _make_family(CVT, "cvt", (2, 4, 6, 7, 8))
globals().update(_fam)
__all__ += [k for k in _fam if not k.startswith("_")]
from .utils.helpers import pe_check  # noqa: F401,E402  (import-path parity with the reference module)
