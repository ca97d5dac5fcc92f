"""B200-native drop-in for the `efficient_attention` package of HKUNLP/efficient-attention.

Same import name and plugin surface as the reference (efficient_attention/__init__.py:1-86):
`AttentionFactory.build_attention`, `AttentionFactory.add_attn_specific_args`,
`add_nested_argument`, `NestedNamespace`, `remove_argument`, plus the module classes.  The modules
own the same parameters under the same names (checkpoints load unchanged) but run the attention core
as hand-written sm_100a CUDA kernels reached through the C ABI of libeva_sm100.so.

Scope: 'eva', 'lara', 'causal_eva', the 'local' / 'softmax' bases they are built on, and (8f-4) the
random-feature baselines of the registry: 'performer', 'ra', 'scatterbrain'.
"""
import argparse
from typing import Dict


def remove_argument(parser, arg):
    """Drop one option (by flag or dest) from an argparse parser (reference __init__.py:5-16)."""
    victim = next((a for a in parser._actions
                   if (a.option_strings and a.option_strings[0] == arg) or a.dest == arg), None)
    if victim is not None:
        parser._remove_action(victim)
    for group in parser._action_groups:
        for action in list(group._group_actions):
            if action.dest == arg:
                group._group_actions.remove(action)
                return


def remove_prefix(text, prefix):
    return text[len(prefix):] if text.startswith(prefix) else text


def add_nested_argument(parser, name, struct_name="attn_args", prefix="", **kwargs):
    """`parser.add_argument` whose dest is `<struct_name>.<option>`; with a prefix, the flag
    `--<prefix>-foo-bar` lands in `<struct_name>.foo_bar` (reference __init__.py:22-27)."""
    stem = name.lstrip('-') if len(prefix) == 0 else remove_prefix(name, "--" + prefix + "-")
    parser.add_argument(name, dest='{}.{}'.format(struct_name, stem.replace('-', '_')), **kwargs)


class NestedNamespace(argparse.Namespace):
    """Namespace that turns dotted attribute names into nested namespaces (reference __init__.py:31-39)."""

    def __setattr__(self, name, value):
        head, dot, rest = name.partition('.')
        if not dot:
            self.__dict__[name] = value
            return
        child = getattr(self, head, None)
        if child is None:
            child = NestedNamespace()
        setattr(child, rest, value)
        self.__dict__[head] = child


from .abstract_attention import MultiheadAttention  # noqa: E402
from .local_attention import LocalAttention  # noqa: E402
from .kernelized_attention import KernelizedAttention  # noqa: E402
from .lara import LinearRA  # noqa: E402
from .randomized_attention import RandomizedAttention  # noqa: E402
from .scatterbrain_attention import ScatterBrain  # noqa: E402
from .eva import EVA  # noqa: E402
from .causal_eva import CausalEVAttention  # noqa: E402
from ._abi import invalidate_caches  # noqa: E402,F401


class AttentionFactory(object):
    attn_dict = {
        'performer': KernelizedAttention,
        'softmax': MultiheadAttention,
        'local': LocalAttention,
        'lara': LinearRA,
        'ra': RandomizedAttention,
        'scatterbrain': ScatterBrain,
        'eva': EVA,
        'causal_eva': CausalEVAttention,
    }

    @classmethod
    def build_attention(cls, attn_name: str, attn_args: Dict):
        # KeyError on an unknown name, TypeError on an unknown kwarg -- as the reference.
        return cls.attn_dict[attn_name](**attn_args)

    @classmethod
    def add_attn_specific_args(cls, parent_parser, attn_name, struct_name="attn_args", prefix=""):
        attn_cls = cls.attn_dict[attn_name]
        if hasattr(attn_cls, 'add_attn_specific_args'):
            return attn_cls.add_attn_specific_args(parent_parser, struct_name=struct_name, prefix=prefix)
        return parent_parser
