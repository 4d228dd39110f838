"""multicore-hw2_b200 -- B200-native brute-force 1-nearest-neighbour path of wu-kan/multicore-hw2.

Host-side mirror of the reference's interface for this path:

* :func:`cudaCallback` -- same name and argument meaning as the reference's entry point
  (/root/reference/sources/src/core.h:71, core.cu:1282-1297 -> v8::cudaCallback 856-958): host
  AoS arrays in, one nearest-reference index per query out.
* :mod:`.device` -- the device-resident building blocks (keys_init / nearest_keys / keys_unpack /
  repack_soa) on torch CUDA tensors; torch only provides memory and streams.
* :mod:`.sharded` -- the multi-GPU path, one process per GPU: contiguous reference shards
  (core.cu:875-883) and one all-reduce(min) of packed uint64 keys over NCCL instead of the
  reference's host-side merge (core.cu:925-957).

Everything computes in ``libnn_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
include/nn_b200.h).  There is no CPU fallback."""
from ._lib import KEY_INIT, LIB_PATH, NNError, build, lib  # noqa: F401
from .api import (Index, cudaCallback, describe_plan, device_count, last_gpus, launch_count, plan_gpus,  # noqa: F401
                  plan_search_groups, probe_fp32, search_host, set_option, shard_range)

__all__ = ["cudaCallback", "search_host", "Index", "describe_plan", "device_count", "launch_count", "plan_gpus", "plan_search_groups", "last_gpus", "set_option",
           "shard_range", "probe_fp32", "KEY_INIT", "LIB_PATH", "NNError", "build", "lib"]
