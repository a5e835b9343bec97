# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference hot path touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.

import torch.distributed as dist
def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
def is_main_process():
    return get_rank() == 0
def synchronize():
    if get_world_size() > 1:
        dist.barrier()
def all_gather(data, group=None):
    return [data]
def gather(data, dst=0, group=None):
    return [data]
