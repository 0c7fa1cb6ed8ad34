"""Runtime switches of the CUDA path (process-wide)."""

# Conservative beam/obstruction culling in the trace kernel.  Culling never changes a ray's result
# (tests/test_gpu_trace.py::test_culling_is_exact compares every ray against brute force); turning
# it off exists for that test and for roofline comparisons.
cull_obstructions = True

# Return NumPy arrays (device -> host copy + synchronisation) instead of torch CUDA tensors from
# render / render_debug / render_response_matrix.  Off by default: outputs stay on the GPU, stream-ordered.
return_numpy = False
