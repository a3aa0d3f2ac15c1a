"""Stand-in for the absent ``optree`` dependency of the reference  --  TEST INFRASTRUCTURE ONLY.

The reference uses exactly one symbol, ``optree.tree_map`` (quantized_tensor.py:561-562), to map a
function over the (args, kwargs) pytree of a torch function call.  torch's own pytree does the same."""
import torch.utils._pytree as _pt


def tree_map(fn, tree, *rest, **kw):
    return _pt.tree_map(fn, tree)
