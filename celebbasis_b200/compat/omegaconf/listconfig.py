"""`omegaconf.listconfig.ListConfig` stand-in (openaimodel.py:476 checks `type(context_dim) == ListConfig`)."""


class ListConfig(list):
    def __deepcopy__(self, memo):
        import copy
        return ListConfig([copy.deepcopy(v, memo) for v in self])
