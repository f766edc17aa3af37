# jBLASB200.jl -- thin `ccall` wrapper over libjblas_b200.so (include/jblas_b200.h).
#
# NOT EXECUTED in the build environment: Julia is not installed there (SURVEY.md s0, s8c).  The executable
# proof of the C ABI is the Python ctypes host in jblas/jl_b200/ (same symbols, same argument order), which the
# GPU tests drive.  This file is what a jBLAS.jl maintainer would add to switch `jmul!` to the B200 path.
#
# Drop-in surface (same names, argument order and column-major conventions as the reference):
#   jmul!(D, A, X [, Val...])   src/gemm.jl:244-348      D = A*X, returns D; the five prefetch Vals are accepted
#                                                         and ignored (async shared-memory staging replaces them)
#   gemm!(D, A, X)              BASELINE.json's name for the same entry
#   fastmul!(D, A, X)           src/kernels.jl:202-208
#   kernel!(pD, pA, pX, K)      src/kernels.jl:239-241   D += A*X on raw pointers
#   initkernel!(pD, pA, pX, K)  src/kernels.jl:273-275   D  = A*X on raw pointers
#   Kernel{Mk,Pk,stride_AD,stride_X,N}   src/kernel_structure.jl:8-9
#   gemm_plus_c!(D, A, X, C)    D = A*X + C    } the fused forms the reference planned but never wrote
#   gemm_x_plus_c!(D, A, X, C)  D = A*(X + C)  } (src/memory_management.jl:72-76)
# Accepted matrix types: anything with `pointer`, `size`, `stride(·,2)` and unit row stride -- MMatrix{M,N,T}
# (the reference's type; `pointer(A)` is a stable dense column-major buffer, src/gemm.jl:309-311) and
# Matrix{T}/StridedMatrix{T}, T in {Float64, Float32}.
module jBLASB200

export jmul!, gemm!, fastmul!, kernel!, initkernel!, Kernel, init, shutdown, gemm_plus_c!, gemm_x_plus_c!, fastmul_batched!

const libjblas_b200 = get(ENV, "JBLAS_B200_LIB", joinpath(@__DIR__, "..", "jblas", "jl_b200", "libjblas_b200.so"))

const F64_AUTO, F64_DMMA, F64_SIMT = Cint(0), Cint(1), Cint(2)
const F32_EXACT, F32_3XTF32 = Cint(0), Cint(1)

struct JblasB200Error <: Exception
    code::Cint
    msg::String
end
Base.showerror(io::IO, e::JblasB200Error) = print(io, "jblas_b200 error ", e.code, ": ", e.msg)

last_error() = unsafe_string(ccall((:jblas_b200_last_error, libjblas_b200), Cstring, ()))
@inline check(rc::Cint) = rc < 0 ? throw(JblasB200Error(rc, last_error())) : rc

"Bind this process to one GPU (one process per GPU). Throws when no CUDA device exists: there is no CPU fallback."
init(device::Integer = 0) = check(ccall((:jblas_b200_init, libjblas_b200), Cint, (Cint,), device))
shutdown() = check(ccall((:jblas_b200_shutdown, libjblas_b200), Cint, ()))

const _initialised = Ref(false)
@inline function ensure_init()
    _initialised[] || (init(parse(Int, get(ENV, "LOCAL_RANK", "0"))); _initialised[] = true)
    nothing
end

# the same singleton the reference defines, src/kernel_structure.jl:8-9
struct Kernel{Mk,Pk,stride_AD,stride_X,N} end
Kernel(Mk, Pk, stride_AD, stride_X, N) = Kernel{Mk,Pk,stride_AD,stride_X,N}()

@inline function _dims(D, A, X)
    M, P = size(D)
    MA, N = size(A)
    NX, PX = size(X)
    (MA == M && NX == N && PX == P) ||
        throw(DimensionMismatch("D is $(M)x$(P), A is $(MA)x$(N), X is $(NX)x$(PX)"))  # MethodError in the reference
    (stride(D, 1) == 1 && stride(A, 1) == 1 && stride(X, 1) == 1) ||
        throw(ArgumentError("matrices must be column-major with unit row stride"))
    M, N, P
end
@inline _ld(A) = size(A, 2) > 1 ? stride(A, 2) : max(size(A, 1), 1)

for (T, gemm, kern, initk) in ((Float64, :jblas_b200_gemm_f64, :jblas_b200_kernel_f64, :jblas_b200_initkernel_f64),
                               (Float32, :jblas_b200_gemm_f32, :jblas_b200_kernel_f32, :jblas_b200_initkernel_f32))
    @eval begin
        # jBLAS naming: D is MxP, A is MxN, X is NxP, N contracted (src/gemm.jl:244).  The C entry is BLAS-named
        # (M, K, N) = (M, N, P).
        function _gemm!(D::AbstractMatrix{$T}, A::AbstractMatrix{$T}, X::AbstractMatrix{$T}, accumulate::Bool, selector::Cint)
            M, N, P = _dims(D, A, X)
            ensure_init()
            GC.@preserve D A X begin  # the reference takes raw pointers without preserving (src/gemm.jl:309-311)
                check(ccall(($(QuoteNode(gemm)), libjblas_b200), Cint,
                            (Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint),
                            pointer(D), pointer(A), pointer(X), M, N, P, _ld(D), _ld(A), _ld(X), accumulate, selector))
            end
            D
        end
        function kernel!(pD::Ptr{$T}, pA::Ptr{$T}, pX::Ptr{$T}, ::Kernel{Mk,Pk,sAD,sX,N}) where {Mk,Pk,sAD,sX,N}
            ensure_init()
            check(ccall(($(QuoteNode(kern)), libjblas_b200), Cint,
                        (Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64), pD, pA, pX, Mk, Pk, sAD, sX, N))
            nothing
        end
        function initkernel!(pD::Ptr{$T}, pA::Ptr{$T}, pX::Ptr{$T}, ::Kernel{Mk,Pk,sAD,sX,N}) where {Mk,Pk,sAD,sX,N}
            ensure_init()
            check(ccall(($(QuoteNode(initk)), libjblas_b200), Cint,
                        (Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64), pD, pA, pX, Mk, Pk, sAD, sX, N))
            nothing
        end
    end
end

for (T, plusc, xplusc) in ((Float64, :jblas_b200_gemm_plus_c_f64, :jblas_b200_gemm_x_plus_c_f64),
                           (Float32, :jblas_b200_gemm_plus_c_f32, :jblas_b200_gemm_x_plus_c_f32))
    for (fn, sym, crows) in ((:_gemm_plus_c!, plusc, :M), (:_gemm_x_plus_c!, xplusc, :N))
        @eval function $fn(D::AbstractMatrix{$T}, A::AbstractMatrix{$T}, X::AbstractMatrix{$T}, C::AbstractMatrix{$T}, selector::Cint)
            M, N, P = _dims(D, A, X)
            (size(C) == ($crows, P) && stride(C, 1) == 1) || throw(DimensionMismatch("C must be $($crows)x$(P), column-major"))
            ensure_init()
            GC.@preserve D A X C begin
                check(ccall(($(QuoteNode(sym)), libjblas_b200), Cint,
                            (Ptr{$T}, Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Cint),
                            pointer(D), pointer(A), pointer(X), pointer(C), M, N, P, _ld(D), _ld(A), _ld(X), _ld(C), selector))
            end
            D
        end
    end
end

_default_selector(::Type{Float64}) = F64_AUTO
_default_selector(::Type{Float32}) = F32_EXACT
_exact_selector(::Type{Float64}) = F64_SIMT
_exact_selector(::Type{Float32}) = F32_EXACT

"""
    jmul!(D, A, X, ::Val=Val(7), ::Val=Val(7), ::Val=Val(3), ::Val=Val(3), ::Val=Val(3); kernel) -> D

`D = A * X` on the B200.  Same signature as jBLAS.jmul! (src/gemm.jl:244-246); the prefetch `Val`s are ignored.
Unlike the reference, remainder rows/columns are computed (src/gemm.jl:266-267,313 skips them).
"""
function jmul!(D::AbstractMatrix{T}, A::AbstractMatrix{T}, X::AbstractMatrix{T},
               ::Val = Val(7), ::Val = Val(7), ::Val = Val(3), ::Val = Val(3), ::Val = Val(3);
               kernel::Cint = _default_selector(T)) where {T<:Union{Float64,Float32}}
    _gemm!(D, A, X, false, kernel)
end
const gemm! = jmul!

"fastmul!(D, A, X) (src/kernels.jl:202-208): exact-chain kernels, any row count."
fastmul!(D::AbstractMatrix{T}, A::AbstractMatrix{T}, X::AbstractMatrix{T}) where {T<:Union{Float64,Float32}} =
    _gemm!(D, A, X, false, _exact_selector(T))

"gemm_plus_c!(D, A, X, C): D = A*X + C; every element's fma chain starts from C[i,j] (kernel!'s accumulate with the start read from C)."
gemm_plus_c!(D::AbstractMatrix{T}, A::AbstractMatrix{T}, X::AbstractMatrix{T}, C::AbstractMatrix{T};
             kernel::Cint = _default_selector(T)) where {T<:Union{Float64,Float32}} = _gemm_plus_c!(D, A, X, C, kernel)
"gemm_x_plus_c!(D, A, X, C): D = A*(X + C), C sized like X; X + C is rounded once per element."
gemm_x_plus_c!(D::AbstractMatrix{T}, A::AbstractMatrix{T}, X::AbstractMatrix{T}, C::AbstractMatrix{T};
               kernel::Cint = _default_selector(T)) where {T<:Union{Float64,Float32}} = _gemm_x_plus_c!(D, A, X, C, kernel)

"""
    fastmul_batched!(D::Array{T,3}, A::Array{T,3}, X::Array{T,3}) -> D

`D[:,:,b] = A[:,:,b] * X[:,:,b]` for every `b` in ONE launch: the B200 form of `fastmul!` (src/kernels.jl:202-208) for a
collection of small matrices (a dense `Array{T,3}` is exactly `batch` column-major matrices back to back).
"""
function fastmul_batched! end

for (T, sym) in ((Float64, :jblas_b200_fastmul_batched_f64), (Float32, :jblas_b200_fastmul_batched_f32))
    @eval function fastmul_batched!(D::Array{$T,3}, A::Array{$T,3}, X::Array{$T,3})
        M, P, B = size(D)
        MA, N, BA = size(A)
        NX, PX, BX = size(X)
        (MA == M && NX == N && PX == P && BA == B && BX == B) || throw(DimensionMismatch("D $(size(D)), A $(size(A)), X $(size(X))"))
        ensure_init()
        GC.@preserve D A X begin
            check(ccall(($(QuoteNode(sym)), libjblas_b200), Cint,
                        (Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64, Int64, Int64),
                        pointer(D), pointer(A), pointer(X), M, N, P, B, M * P, M * N, N * P))
        end
        D
    end
end

end # module
