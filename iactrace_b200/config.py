"""Runtime switches of the CUDA path (process-wide)."""
import os as _os

# Conservative beam/obstruction culling in the trace kernel.  Culling never changes a ray's result
# (tests/test_gpu_trace.py::test_culling_is_exact compares every ray against brute force); turning
# it off exists for that test and for roofline comparisons.
cull_obstructions = True

# Return NumPy arrays (device -> host copy + synchronisation) instead of torch CUDA tensors from
# render / render_debug / render_response_matrix.  Off by default: outputs stay on the GPU, stream-ordered.
return_numpy = False

# From this many samples per facet on, the world table is written in spatially binned order
# (iact_transform_to_world_binned) and the trace kernel culls again per 32-sample run.  Below it the
# runs are too large a part of the facet to shed candidates (CT5, M = 115: 1.56 -> 1.20 cylinder tests per
# ray, 4.7 -> 5.2 ms).  0 disables binning.  The environment overrides are for tuning runs.
bin_samples_min = int(_os.environ.get("IACTRACE_B200_BIN_SAMPLES_MIN", "256"))
# ... and only for scenes with at least this many obstruction primitives: with a handful of them the
# candidate lists are mostly empty.  (CT3, 33 cylinders, M = 1000: 12.9 -> 11.8 ms with the strip test.)
bin_obstructions_min = int(_os.environ.get("IACTRACE_B200_BIN_OBSTRUCTIONS_MIN", "16"))

# Large sample counts.  A world table that does not fit the 126 MB L2 is re-read from HBM once per source (the work
# queue is source-major), so `render` / `render_response_matrix` walk the samples of every facet in windows whose world
# table (F x window x 32 B) stays below `window_table_bytes`, summing the window images.  Groups sampled with
# MCIntegrator(n_samples) whose tables would exceed `stream_samples_bytes` are not materialised at all: each window is
# regenerated from the counter-based stream (core/integrators.py SampleStream), bit-identical to the full draw.
window_table_bytes = int(_os.environ.get("IACTRACE_B200_WINDOW_TABLE_BYTES", str(48 << 20)))
stream_samples_bytes = int(_os.environ.get("IACTRACE_B200_STREAM_SAMPLES_BYTES", str(1 << 30)))
