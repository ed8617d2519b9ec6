"""B200-native serving hot path of the two-tower recommender (HeikalPro/two-tower-model-v2).

Drop-in replacements for the reference's two hot-path modules, backed by hand-written sm_100a
CUDA kernels behind the C-ABI in include/tt_b200.h:

    reference                              here
    src/models/buyer_tower.BuyerTower      two_tower_model_v2_b200.BuyerTower
    src/inference/vector_db.VectorDatabase two_tower_model_v2_b200.VectorDatabase
    src/training/losses.InfoNCELoss        two_tower_model_v2_b200.InfoNCELoss

The directory is named `two-tower-model-v2_b200`; import it as `two_tower_model_v2_b200`
(the root-level `two_tower_model_v2_b200.py` aliases it).
"""
from .batcher import ArrayRows, MicroBatcher
from .buyer_tower import BuyerTower
from .config import get_event_weight
from .losses import InfoNCELoss
from .retrieval import RetrievalPipeline, ShardedRetrievalPipeline
from .sharded import ShardedFlatIPIndex, shard_bounds
from .vector_db import (FlatIPIndex, VectorDatabase, read_flat_ip_file, read_native_shard, write_flat_ip_file,
                        write_native_shard)

__all__ = ["BuyerTower", "InfoNCELoss", "VectorDatabase", "FlatIPIndex", "ShardedFlatIPIndex", "RetrievalPipeline", "ShardedRetrievalPipeline", "MicroBatcher", "ArrayRows", "shard_bounds",
           "get_event_weight", "read_flat_ip_file", "write_flat_ip_file", "read_native_shard", "write_native_shard"]
