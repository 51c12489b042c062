"""B200 data plane: flat arenas, symmetric (peer-mapped) memory, fused reduce+optimizer kernels,
CUDA-graph step capture and the NVLink learners."""
