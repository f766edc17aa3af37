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
#   kernel!/initkernel!(pD, pA, pX, K, pf::PrefetchA|PrefetchX|PrefetchAX)   src/kernels.jl:277-537 (pf ignored)
#   PrefetchA, PrefetchX, PrefetchAX     src/memory_management.jl:3-12
#   prefetch(address, Val, Val)          src/memory_management.jl:45-54  (exported by the reference; a no-op here)
#   mrandn(M, N), mrandn(T, M, N)        src/randmat.jl:11-14  (exported by the reference)
#   jmul!(D, A, X; gpus = n)             the single-node multi-GPU mode (north_star (3)): jblas_b200_mgpu_gemm_*
#   gemm_plus_c!(D, A, X, C)    D = A*X + C    } the fused forms the reference planned but never wrote
#   gemm_x_plus_c!(D, A, X, C)  D = A*(X + C)  } (src/memory_management.jl:72-76)
# Accepted matrix types: anything with `pointer`, `size`, `stride(·,2)` and unit row stride -- MMatrix{M,N,T}
# (the reference's type; `pointer(A)` is a stable dense column-major buffer, src/gemm.jl:309-311) and
# Matrix{T}/StridedMatrix{T}, T in {Float64, Float32}.
module jBLASB200

# the reference's own exports (src/jBLAS.jl:6-8) ...
export mrandn, jmul!, prefetch
# ... its unexported-but-used names (SURVEY.md s1), and what this path adds
export gemm!, fastmul!, kernel!, initkernel!, Kernel, PrefetchA, PrefetchX, PrefetchAX, init, shutdown, mgpu_init, pin!, unpin!,
       gemm_plus_c!, gemm_x_plus_c!, fastmul_batched!

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

"""
    mgpu_init(ngpus = 0) -> Int

Create the per-GPU contexts of the single-process multi-GPU mode on devices `0:ngpus-1` and enable peer access
(`ngpus <= 0`: every visible GPU).  `jmul!(D, A, X; gpus = n)` calls it on demand.
"""
function mgpu_init(ngpus::Integer = 0)
    n = check(ccall((:jblas_b200_mgpu_init, libjblas_b200), Cint, (Cint,), ngpus))
    _initialised[] = true  # the library bound the process to device 0 if it was not bound yet
    Int(n)
end

"""
    pin!(A) -> A;  unpin!(A) -> A

Page-lock (`jblas_b200_host_register`) / release a caller-owned array so the host-pointer entries DMA at full PCIe rate.
Optional: unpinned arrays work, only slower.  Pin once per buffer, not once per call.
"""
function pin!(A::Union{DenseArray{T},Base.ReshapedArray{T}}) where {T}
    ensure_init()
    check(ccall((:jblas_b200_host_register, libjblas_b200), Cint, (Ptr{Cvoid}, Csize_t), pointer(A), sizeof(T) * length(A)))
    A
end
function unpin!(A)
    check(ccall((:jblas_b200_host_unregister, libjblas_b200), Cint, (Ptr{Cvoid},), pointer(A)))
    A
end

# the same singleton the reference defines, src/kernel_structure.jl:8-9
struct Kernel{Mk,Pk,stride_AD,stride_X,N} end
Kernel(Mk, Pk, stride_AD, stride_X, N) = Kernel{Mk,Pk,stride_AD,stride_X,N}()

# Prefetch offsets of the reference's tile kernels (src/memory_management.jl:3-12): byte offsets the CPU kernels add to
# pA / pX when issuing software prefetches (src/kernels.jl:277-537).  Kept so that code written against the reference
# compiles and runs unchanged; on the B200 path the TMA / cp.async shared-memory rings do the staging, so the values
# are accepted and ignored.
struct PrefetchA
    A::Int
end
struct PrefetchX
    X::Int
end
struct PrefetchAX
    A::Int
    X::Int
end

"""
    prefetch(address, Val(Locality) = Val(1), Val(ReadOrWrite) = Val(0)) -> nothing

The reference's exported software prefetch (`llvm.prefetch`, src/memory_management.jl:45-54).  A no-op here: host
addresses are not what the GPU reads, and the kernels stage their tiles asynchronously on their own (a10 in SURVEY.md s8).
"""
@inline prefetch(address, ::Val = Val(1), ::Val = Val(0)) = nothing

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

# The six prefetch variants (src/kernels.jl:277, 320, 378, 407, 444, 485): same product, `pf` accepted and ignored.
for PF in (:PrefetchAX, :PrefetchA, :PrefetchX)
    @eval begin
        kernel!(pD::Ptr{T}, pA::Ptr{T}, pX::Ptr{T}, K::Kernel, ::$PF) where {T<:Union{Float64,Float32}} = kernel!(pD, pA, pX, K)
        initkernel!(pD::Ptr{T}, pA::Ptr{T}, pX::Ptr{T}, K::Kernel, ::$PF) where {T<:Union{Float64,Float32}} = initkernel!(pD, pA, pX, K)
    end
end

# single-process multi-GPU form of _gemm! (jblas_b200_mgpu_gemm_*): column blocks of X and D per GPU, src/gemm.jl:313
for (T, sym) in ((Float64, :jblas_b200_mgpu_gemm_f64), (Float32, :jblas_b200_mgpu_gemm_f32))
    @eval function _mgpu_gemm!(D::AbstractMatrix{$T}, A::AbstractMatrix{$T}, X::AbstractMatrix{$T}, accumulate::Bool, selector::Cint, gpus::Integer)
        M, N, P = _dims(D, A, X)
        GC.@preserve D A X begin
            check(ccall(($(QuoteNode(sym)), libjblas_b200), Cint,
                        (Ptr{$T}, Ptr{$T}, Ptr{$T}, Int64, Int64, Int64, Int64, Int64, Int64, Cint, Cint, Cint),
                        pointer(D), pointer(A), pointer(X), M, N, P, _ld(D), _ld(A), _ld(X), accumulate, selector, gpus))
        end
        _initialised[] = true
        D
    end
end

# mrandn (src/randmat.jl:11-14): iid N(0,1), generated on the GPU (Philox, seeded) and copied back.  The reference returns an
# MMatrix{M,N,Float64}; StaticArrays is not a dependency of this wrapper, so a Matrix{T} comes back (pass it to
# `MMatrix{M,N}(...)` if the static type is wanted; SURVEY.md App. A: large MMatrix types compile slowly anyway).
function mrandn(::Type{T}, M::Integer, N::Integer; seed::Integer = 0x6a424c41) where {T<:Union{Float64,Float32}}
    ensure_init()
    out = Matrix{T}(undef, M, N)
    bytes = sizeof(T) * M * N
    bytes == 0 && return out
    dptr = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jblas_b200_alloc, libjblas_b200), Cint, (Ptr{Ptr{Cvoid}}, Csize_t), dptr, bytes))
    try
        check(ccall((:jblas_b200_randn_fill, libjblas_b200), Cint, (Ptr{Cvoid}, Int64, Int64, UInt64, Cint, Ptr{Cvoid}),
                    dptr[], 0, M * N, UInt64(seed), T === Float64 ? Cint(0) : Cint(1), C_NULL))
        GC.@preserve out check(ccall((:jblas_b200_d2h, libjblas_b200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), pointer(out), dptr[], bytes))
    finally
        ccall((:jblas_b200_free, libjblas_b200), Cint, (Ptr{Cvoid},), dptr[])
    end
    out
end
mrandn(M::Integer, N::Integer; kw...) = mrandn(Float64, M, N; kw...)

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
    jmul!(D, A, X, ::Val=Val(7), ::Val=Val(7), ::Val=Val(3), ::Val=Val(3), ::Val=Val(3); kernel, gpus = 1) -> D

`D = A * X` on the B200.  Same signature as jBLAS.jmul! (src/gemm.jl:244-246); the prefetch `Val`s are ignored.
Unlike the reference, remainder rows/columns are computed (src/gemm.jl:266-267,313 skips them).
"""
function jmul!(D::AbstractMatrix{T}, A::AbstractMatrix{T}, X::AbstractMatrix{T},
               ::Val = Val(7), ::Val = Val(7), ::Val = Val(3), ::Val = Val(3), ::Val = Val(3);
               kernel::Cint = _default_selector(T), gpus::Integer = 1) where {T<:Union{Float64,Float32}}
    # gpus > 1: the single-node multi-GPU mode -- GPU g owns a column block of X and D (the outer column-tile loop of the
    # reference, src/gemm.jl:313), A reaches every GPU in K panels; bit-identical to the one-GPU result
    gpus == 1 ? _gemm!(D, A, X, false, kernel) : _mgpu_gemm!(D, A, X, false, kernel, gpus)
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
