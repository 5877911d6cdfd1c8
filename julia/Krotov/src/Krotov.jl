# Krotov -- drop-in replacement of JuliaQuantumControl/Krotov.jl whose hot path runs in libkrotov_cuda (B200, sm_100a).
#
# NOT EXECUTED IN THIS REPOSITORY: Julia is not installed in the build image or on the GPU boxes.  The C ABI these
# files bind (include/krotov_cuda.h) is exercised through the Python mirror (krotov.jl_b200/), and
# tests/test_host.py checks the struct layouts and ccall signatures below against the header, so that the two cannot
# drift apart.  Behaviour follows the reference's driver (src/optimize.jl:155-235, 374-496), workspace
# (src/workspace.jl:65-200) and result type (src/result.jl:34-104); the two hot functions
# `krotov_initial_fw_prop!` and `krotov_iteration` are one `ccall` each.
module Krotov

include("LibKrotovCuda.jl")
include("result.jl")
include("workspace.jl")
include("optimize.jl")

end
