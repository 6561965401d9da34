"""Run-time switches of the drop-in layer."""

# Raise the reference's IndexError when an event indexes outside the output (costs one 8-byte
# device->host read per call).  With False, out-of-range events are silently dropped.
check_index_errors = True

# Kernel variant forced for the scatter entry points (None = library heuristic).
# One of None, "global_red", "vector_red", "warp_agg".
variant = None

# Device used when the caller hands over host (CPU / numpy) arrays.
default_device = "cuda"
