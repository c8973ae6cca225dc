"""gmeta_b200: B200-native (sm_100a) implementation of the G-Meta inner-loop hot path.

Drop-in surface of the reference (mims-harvard/G-Meta):
    from gmeta_b200.meta import Meta, proto_loss_spt, proto_loss_qry, euclidean_dist
    from gmeta_b200.learner import Classifier
The arithmetic lives in libgmeta_b200.so (C ABI: include/gmeta_b200.h); there is no CPU path.
"""
from ._lib import GMetaError, build_library, lib  # noqa: F401
from .packed import PackedSubgraphBatch, SubgraphCSR  # noqa: F401

__all__ = ["GMetaError", "build_library", "lib", "PackedSubgraphBatch", "SubgraphCSR"]
