# make_reference_golden.jl -- PIN the parity of this repository to the REAL jBLAS.jl.
#
# Why this exists.  The build environment of jblas-b200 has no Julia, and jBLAS.jl's SIMD dependencies (SIMDPirates 0.1.0,
# VectorizationBase 0.1.0) are un-vendored path-dev checkouts (Manifest.toml:79-83, 120-124), so the reference could never be
# run there.  The goldens under tests/golden/*.npz were therefore produced by a line-by-line emulation of the code `jmul!`
# generates (oracle/structural_jmul.py) and parity is labelled "unpinned" (DESIGN.md s2).  Anyone with a Julia in which
# `using jBLAS` works can replace them with the reference's own output:
#
#     julia --project=/path/to/jBLAS.jl julia/make_reference_golden.jl tests/golden
#     python -m pytest tests/test_oracle.py -q            # the oracle must still reproduce every D bit for bit
#
# What it does: for every tests/golden/<name>.npz it reads A and X (little-endian .npy members, column-major), calls the
# real `jBLAS.jmul!` (src/gemm.jl:244), `jBLAS.fastmul!` (src/kernels.jl:202) or `jBLAS.kernel!`/`initkernel!`
# (src/kernels.jl:239,273) as named by the file's `call` member (default: jmul!), and rewrites the `D` member and a
# `covered` member (rows, cols the reference's full-tile loops write: it skips remainder rows/columns, src/gemm.jl:266-267).
# Elements outside `covered` keep the NaN sentinel.  The .npz container is a plain zip of .npy files; both are parsed here
# without any Python.
#
# This script has NOT been executed (no Julia in the build environment).  It uses only Base, ZipFile.jl and StaticArrays.
using jBLAS, StaticArrays
import ZipFile

function read_npy(io::IO)
    magic = read(io, 6)
    magic == UInt8[0x93, 0x4e, 0x55, 0x4d, 0x50, 0x59] || error("not an .npy stream")
    major = read(io, UInt8); read(io, UInt8)
    hlen = major == 1 ? Int(ltoh(read(io, UInt16))) : Int(ltoh(read(io, UInt32)))
    header = String(read(io, hlen))
    descr = match(r"'descr':\s*'([^']+)'", header).captures[1]
    fortran = match(r"'fortran_order':\s*(True|False)", header).captures[1] == "True"
    shape = Tuple(parse.(Int, filter(!isempty, split(match(r"'shape':\s*\(([^)]*)\)", header).captures[1], r"[,\s]+"))))
    T = descr == "<f8" ? Float64 : descr == "<f4" ? Float32 : descr == "<i8" ? Int64 : error("dtype $descr not handled")
    data = Vector{T}(undef, prod(shape))
    read!(io, data)
    data .= ltoh.(data)
    length(shape) <= 1 && return data
    fortran ? reshape(data, shape) : permutedims(reshape(data, reverse(shape)), reverse(1:length(shape)))
end

function write_npy(io::IO, A::AbstractArray{T}) where {T}
    descr = T === Float64 ? "<f8" : T === Float32 ? "<f4" : T === Int64 ? "<i8" : error("dtype")
    shape = join(size(A), ", ") * (ndims(A) == 1 ? "," : "")
    header = "{'descr': '$descr', 'fortran_order': True, 'shape': ($shape), }"
    pad = 64 - (10 + length(header) + 1) % 64
    header *= " "^(pad % 64) * "\n"
    write(io, UInt8[0x93, 0x4e, 0x55, 0x4d, 0x50, 0x59, 0x01, 0x00])
    write(io, htol(UInt16(length(header))))
    write(io, header)
    write(io, htol.(vec(collect(A))))
end

function regenerate(path::AbstractString)
    zr = ZipFile.Reader(path)
    members = Dict{String,Any}()
    for f in zr.files
        members[replace(f.name, ".npy" => "")] = read_npy(f)
    end
    close(zr)
    A, X = members["A"], members["X"]
    T = eltype(A)
    M, N = size(A)
    P = size(X, 2)
    D = fill(T(NaN), M, P)
    mA, mX, mD = MMatrix{M,N,T}(A), MMatrix{N,P,T}(X), MMatrix{M,P,T}(D)
    call = haskey(members, "call") ? String(Char.(members["call"])) : "jmul!"
    if call == "fastmul!"
        jBLAS.fastmul!(mD, mA, mX)              # src/kernels.jl:202-208: every row (masked remainder), every column
        covered = (M, P)
    else
        jBLAS.jmul!(mD, mA, mX)                 # src/gemm.jl:244-348: full tiles only
        _, rows, cols = jBLAS.pick_kernel_size(T)   # src/kernel_structure.jl:76-99
        covered = ((M ÷ rows) * rows, (P ÷ cols) * cols)
    end
    members["D"] = Array(mD)
    members["covered"] = Int64[covered...]
    members["source"] = Float64[1.0]            # 1.0 = produced by the real jBLAS.jl (0.0 / absent = structural emulation)
    zw = ZipFile.Writer(path)
    for (name, arr) in members
        f = ZipFile.addfile(zw, name * ".npy"; method = ZipFile.Deflate)
        write_npy(f, arr)
    end
    close(zw)
    println("re-pinned ", path, "  (", call, ", covered ", covered, ")")
end

dir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden")
for f in sort(readdir(dir))
    endswith(f, ".npz") && regenerate(joinpath(dir, f))
end
