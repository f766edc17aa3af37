"""jblas -- namespace package; the B200-native jBLAS.jl hot path lives in `jblas.jl_b200`."""
