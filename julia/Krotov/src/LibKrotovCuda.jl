# ccall bindings for libkrotov_cuda (include/krotov_cuda.h).  NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not
# installed in the build image; the same ABI is exercised through Python ctypes (krotov.jl_b200/_lib.py).
module LibKrotovCuda

const lib = get(ENV, "LIBKROTOV_CUDA", "libkrotov_cuda")

const KROTOV_GEN_DENSE_COLMAJOR = Cint(0)
const KROTOV_FORWARD, KROTOV_BACKWARD = Cint(0), Cint(1)
const KROTOV_CHI_HOST, KROTOV_CHI_SM, KROTOV_CHI_SS, KROTOV_CHI_RE = Cint(0), Cint(1), Cint(2), Cint(3)

# mirrors `krotov_problem` field by field
struct Problem
    struct_size::Int32; d::Int32; n_traj::Int32; n_ctrl::Int32; n_steps::Int32; n_gen::Int32
    gen_format::Int32; nnz::Int32
    tlist::Ptr{Float64}; gen_of_traj::Ptr{Int32}; csr_rowptr::Ptr{Int32}; csr_colind::Ptr{Int32}
    gen_values::Ptr{ComplexF64}; term_present::Ptr{UInt8}; psi0::Ptr{ComplexF64}; target::Ptr{ComplexF64}
    weight::Ptr{Float64}; update_shape::Ptr{Float64}; lambda_a::Ptr{Float64}
    functional::Int32; n_traj_global::Int32; store_fw::Int32; device::Int32; force_path::Int32
    reserved::NTuple{7,Int32}
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(ptr)
        h = new(ptr)
        finalizer(h -> (h.ptr != C_NULL && ccall((:krotov_destroy, lib), Cint, (Ptr{Cvoid},), h.ptr); h.ptr = C_NULL), h)
    end
end

last_error(h) = unsafe_string(ccall((:krotov_last_error, lib), Cstring, (Ptr{Cvoid},), h === nothing ? C_NULL : h.ptr))
check(h, rc) = rc == 0 ? nothing : error("libkrotov_cuda: " * last_error(h))   # ErrorException, caught by the
                                                                                # try block of optimize_krotov

function create(p::Problem)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:krotov_create, lib), Cint, (Ref{Problem}, Ref{Ptr{Cvoid}}), p, out)
    rc == 0 || error("libkrotov_cuda: " * last_error(nothing))
    Handle(out[])
end

set_cheby(h, dir, dtc_of_step::Vector{Int32}, dt_of_class::Vector{Float64}, E_min::Vector{Float64},
          Delta::Vector{Float64}, m::Matrix{Int32}, coeffs::Array{Float64,3}) =   # m: [ndtc, n_gen], coeffs: [m_max, ndtc, n_gen]
    check(h, ccall((:krotov_set_cheby, lib), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Cint),
        h.ptr, dir, length(dt_of_class), dtc_of_step, dt_of_class, E_min, Delta, m, coeffs, size(coeffs, 1)))

forward(h, pulses::Matrix{Float64}) =   # pulses: [N_T, L] column-major == [L][N_T] row-major on the wire
    check(h, ccall((:krotov_forward, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, pulses))

set_chi(h, chi::Matrix{ComplexF64}) =   # [d, N]
    check(h, ccall((:krotov_set_chi, lib), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), h.ptr, chi))

function iterate!(h, guess::Matrix{Float64}, new::Matrix{Float64}, g_a_int::Vector{Float64})
    check(h, ccall((:krotov_iterate, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   h.ptr, guess, new, g_a_int))
end

get_states!(h, states::Matrix{ComplexF64}) =
    check(h, ccall((:krotov_get_states, lib), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), h.ptr, states))
get_tau!(h, tau::Vector{ComplexF64}) =
    check(h, ccall((:krotov_get_tau, lib), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), h.ptr, tau))
get_storage!(h, which, k, n0, n1, out::Matrix{ComplexF64}) =
    check(h, ccall((:krotov_get_storage, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{ComplexF64}),
                   h.ptr, which, k, n0, n1, out))

# chi(T) = coef[k] * target[k] formed on the device (built-in functionals on several ranks: the global sum of tau)
set_chi_coeffs(h, coef::Vector{ComplexF64}) =
    check(h, ccall((:krotov_set_chi_coeffs, lib), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), h.ptr, coef))

# Multi-GPU (one Julia process per GPU): exchange the descriptors by any transport (MPI.Allgather, Distributed ...),
# then connect; from then on `iterate!` exchanges the per-time-step overlap sums in-kernel over NVLink.
const COMM_DESC_BYTES = 192
function comm_export(h)
    desc = Vector{UInt8}(undef, COMM_DESC_BYTES)
    check(h, ccall((:krotov_comm_export, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h.ptr, desc))
    desc
end
comm_connect(h, rank::Integer, world::Integer, descs::Matrix{UInt8}) =   # descs: [COMM_DESC_BYTES, world]
    check(h, ccall((:krotov_comm_connect, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), h.ptr, rank, world, descs))

# Optional host utility: the spectral envelope of every generator of an ensemble in one threaded call
# (H0: [d, d, n_gen], Hc: [d, d, n_gen, L], amps: [L, n_corner]); `eigvals` works unchanged.
function envelope_extremes(H0::Array{ComplexF64,3}, Hc::Array{ComplexF64,4}, amps::Matrix{Float64}; threads = 0)
    d, n_gen, L = size(H0, 1), size(H0, 3), size(Hc, 4)
    e_min, e_max = Vector{Float64}(undef, n_gen), Vector{Float64}(undef, n_gen)
    rc = ccall((:krotov_envelope_extremes, lib), Cint,
        (Cint, Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
        n_gen, d, L, H0, Hc, size(amps, 2), amps, e_min, e_max, threads)
    rc == 0 || error("libkrotov_cuda: krotov_envelope_extremes: bad argument")
    e_min, e_max
end

end # module
