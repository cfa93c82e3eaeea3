"""repconc_b200 -- B200-native (sm_100a) implementation of the RepCONC constrained-clustering
product-quantization hot path, behind the reference's own module / function interface.

    from repconc_b200 import RepCONC                     # models/repconc/modeling_repconc.py
    from repconc_b200.evaluate_repconc import (          # models/repconc/evaluate_repconc.py:78-206
        initialize_index, add_docs, from_pq_to_ivfpq, load_index_to_gpu, search, batch_search)

All arithmetic runs in librepconc_b200.so (build: `python -m repconc_b200.build`); importing the
package does not need a GPU, calling into it does.
"""
from .modeling_repconc import RepCONC, QuantizeOutput, decode, sinkhorn_algorithm  # noqa: F401

__version__ = "0.1.0"
