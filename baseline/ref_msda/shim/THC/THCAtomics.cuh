// Build shim for baseline/ref_msda (intentionally empty): float / double atomicAdd are native on sm_100a.
