"""Config-dict construction helpers with the reference's semantics.

Mirrors the interface of the reference's ``src/experiment/config_helpers.py`` (ConfigMixin :6-16,
LazyConstructor :32-47, construct_from_module :50-56, lazy_construct_from_module :59-71): a nested dict
``{'name': ClassName, 'kwargs': {...}}`` is resolved against a module; kernels are resolved *lazily* because their
constructor needs the input dimension, which is only known at fit time (core_models.py:192).
"""


class ConfigMixin:
    """``Class.from_config(kwargs)`` == ``Class(**Class.process_config(**kwargs))`` plus optional overrides."""

    @classmethod
    def from_config(cls, kwargs, processed_kwargs=None):
        merged = dict(cls.process_config(**kwargs))
        if processed_kwargs:
            merged.update(processed_kwargs)
        return cls(**merged)

    @classmethod
    def process_config(cls, **kwargs):
        return kwargs


class LazyConstructor:
    """Holds a class and default kwargs; calling it constructs the object (call-site kwargs win)."""

    def __init__(self, class_, **default_kwargs):
        self.class_ = class_
        self.default_kwargs = default_kwargs

    def __call__(self, *args, **new_kwargs):
        kwargs = dict(self.default_kwargs)
        kwargs.update(new_kwargs)
        return self.class_(*args, **kwargs)


def _split(config, overrides):
    kwargs = dict(config.get('kwargs', {}))
    if overrides:
        kwargs.update(overrides)
    return config['name'], kwargs


def construct_from_module(module, config, overrides=None):
    if config is None:
        return None
    name = config['name']
    return getattr(module, name).from_config(config.get('kwargs', {}), overrides)


def lazy_construct_from_module(module, config, overrides=None):
    if config is None:
        return None
    name, kwargs = _split(config, overrides)
    return LazyConstructor(getattr(module, name), **kwargs)
