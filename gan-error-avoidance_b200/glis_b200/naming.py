"""Reference-compatible child names.

The reference registers children under dotted names such as ``'level.0.conv'`` and
``'lis.0-1.linear'`` (common/model.py:32,182); its checkpoints therefore carry keys like
``level.0.conv.weight``.  Modern torch refuses dots in child names, so
children are registered with ``SEP`` in place of each dot and the keys are translated on
``state_dict()`` / ``load_state_dict()``.
"""
import torch.nn as nn

SEP = "·"


class DottedSequential(nn.Sequential):
    def __init__(self):
        super(DottedSequential, self).__init__()
        self._register_state_dict_hook(DottedSequential._to_dotted)
        self._register_load_state_dict_pre_hook(self._from_dotted)

    def add_module(self, name, module):
        super(DottedSequential, self).add_module(name.replace(".", SEP), module)

    def child(self, dotted_name):
        return self._modules[dotted_name.replace(".", SEP)]

    def named_dotted_children(self):
        for name, m in self._modules.items():
            yield name.replace(SEP, "."), m

    @staticmethod
    def _to_dotted(module, state, prefix, _local_metadata):
        start = len(prefix)
        for key in [k for k in state if k.startswith(prefix) and SEP in k[start:]]:
            state[prefix + key[start:].replace(SEP, ".")] = state.pop(key)

    def _from_dotted(self, state, prefix, *_unused):
        for name in self._modules:
            if SEP not in name:
                continue
            dotted = prefix + name.replace(SEP, ".") + "."
            for key in [k for k in state if k.startswith(dotted)]:
                state[prefix + name + "." + key[len(dotted):]] = state.pop(key)
