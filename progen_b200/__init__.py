"""progen_b200 — B200-native ProGen training + sampling engine (drop-in for lucidrains/progen's `ProGen`)."""


def __getattr__(name):
    if name == 'ProGen':
        from .progen import ProGen
        return ProGen
    raise AttributeError(name)
