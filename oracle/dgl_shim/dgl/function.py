"""ORACLE / TEST INFRASTRUCTURE ONLY.  `dgl.function` stand-in (learner.py:3,9,38-39)."""


class _Fn(object):
    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)


def copy_src(src, out):
    return _Fn("copy_src", src=src, out=out)


def sum(msg, out):  # noqa: A001  (name fixed by the DGL API)
    return _Fn("sum", msg=msg, out=out)
