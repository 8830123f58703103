"""diffusers.configuration_utils.FrozenDict (configuration_utils.py:52-84): the config container a reference-written module pickle
carries in `_internal_dict`.  Read-only OrderedDict whose keys are also attributes."""
from collections import OrderedDict


class FrozenDict(OrderedDict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in self.items():
            object.__setattr__(self, k, v)
        object.__setattr__(self, "_FrozenDict__frozen", True)

    def __getattr__(self, k):          # instances rebuilt by pickle get their items after __init__
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __delitem__(self, *a, **k):
        raise Exception(f"You cannot use ``__delitem__`` on a {self.__class__.__name__} instance.")

    def pop(self, *a, **k):
        raise Exception(f"You cannot use ``pop`` on a {self.__class__.__name__} instance.")

    def update(self, *a, **k):
        raise Exception(f"You cannot use ``update`` on a {self.__class__.__name__} instance.")

    def setdefault(self, *a, **k):
        raise Exception(f"You cannot use ``setdefault`` on a {self.__class__.__name__} instance.")

    def __setattr__(self, name, value):
        if self.__dict__.get("_FrozenDict__frozen", False):
            raise Exception(f"You cannot use ``__setattr__`` on a {self.__class__.__name__} instance.")
        super().__setattr__(name, value)

    def __setitem__(self, name, value):
        if self.__dict__.get("_FrozenDict__frozen", False):
            raise Exception(f"You cannot use ``__setattr__`` on a {self.__class__.__name__} instance.")
        super().__setitem__(name, value)
