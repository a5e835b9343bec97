// Build shim for baseline/ref_msda (intentionally empty): the reference's kernel header includes ATen but uses none of it.
