"""Runtime switches of the CUDA path (process-wide)."""

# Conservative beam/obstruction culling in the trace kernel.  Culling never changes a ray's result
# (tests/test_gpu_trace.py::test_culling_is_exact compares every ray against brute force); turning
# it off exists for that test and for roofline comparisons.
cull_obstructions = True

# Return NumPy arrays (device -> host copy + synchronisation) instead of torch CUDA tensors from
# render / render_debug / render_response_matrix.  Off by default: outputs stay on the GPU, stream-ordered.
return_numpy = False

# From this many samples per facet on, the world table is written in spatially binned order
# (iact_transform_to_world_binned) and the trace kernel culls again per 32-sample run.  Below it the
# per-iteration test costs more than it saves.  0 disables binning.
bin_samples_min = 256
# ... and only for scenes with at least this many obstruction primitives: with few of them the
# candidate lists are mostly empty and the per-run test does not pay (CT3: 33 cylinders, -5 %).
bin_obstructions_min = 100
