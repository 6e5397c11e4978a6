"""jvector_b200 — B200 (sm_100a) implementation of JVector's distance / quantised-scoring hot path.

The product is the C-ABI library `lib/libjvector_b200.so` (include/jvector_b200.h). This package is the host-side
mirror of the reference's interface for that path, used by the tests and the bench:

  VectorSimilarityFunction                      base:vector/VectorSimilarityFunction.java
  F32Vectors / PQVectors / BQVectors / NVQVectors   RandomAccessVectorValues / CompressedVectors (device resident)
  ScoreFunction (similarityTo / similarityToBatch)  base:graph/similarity/ScoreFunction.java
  GraphIndex, GraphSearcher.search              base:graph/GraphSearcher.java:222
  GraphIndexBuilder.build                       base:graph/GraphIndexBuilder.java:436
"""
from . import api
from .api import (BQVectors, F32Vectors, GraphIndex, GraphIndexBuilder, GraphSearcher, NVQVectors, PQVectors, QueryBatch, ScoreFunction,
                  SearchResult, VectorSimilarityFunction, bq_encode_all, kmeans_assign, nvq_encode_all, nvq_encode_resident, pack_accept_bits, pq_encode_all,
                  score_multi, topk_bruteforce)
from ._native import JVectorB200Error, init, load

__all__ = ["VectorSimilarityFunction", "F32Vectors", "PQVectors", "BQVectors", "NVQVectors", "ScoreFunction", "GraphIndex",
           "GraphSearcher", "GraphIndexBuilder", "SearchResult", "score_multi", "topk_bruteforce", "bq_encode_all",
           "pq_encode_all", "nvq_encode_all", "nvq_encode_resident", "QueryBatch", "kmeans_assign", "pack_accept_bits", "JVectorB200Error", "init", "load"]
