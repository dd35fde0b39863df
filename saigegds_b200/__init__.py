"""saigegds_b200: B200-native GRM-vector product and null-model fit behind SAIGEgds' own native interface.

Only the hot path of seqFitNullGLMM_SPA is here: packed-genotype store, GRM product,
PCG, trace / AI-REML / variance-ratio drivers.  All numerics run in libsaigegds_b200.so (CUDA, sm_100a).
"""
from ._lib import InvalidArgument, OverflowErrorSGB, SgbError, build  # noqa: F401
from .api import (Context, DeviceArray, NullModel, default_context, make_param, saige_get_sparse,  # noqa: F401
                  seqFitNullGLMM_SPA, sparse_to_packed)
from .assoc import ScoreTest, init_nullmod, seqAssocGLMM_SPA  # noqa: F401
from .dist import init_comm_from_torch, shard_range  # noqa: F401
from .gds import GdsFormatError, read_gds_genotypes, store_from_gds  # noqa: F401

__all__ = ["Context", "DeviceArray", "NullModel", "default_context", "make_param", "seqFitNullGLMM_SPA", "saige_get_sparse", "sparse_to_packed", "ScoreTest", "init_nullmod", "seqAssocGLMM_SPA",
           "init_comm_from_torch", "shard_range", "read_gds_genotypes", "store_from_gds", "GdsFormatError", "build", "SgbError", "InvalidArgument", "OverflowErrorSGB"]
