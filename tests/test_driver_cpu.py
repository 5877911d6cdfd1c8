"""The driver-semantics tests of tests/test_parity_gpu.py (callbacks, records, continue_from, atexit, storages ... --
the behaviour of the reference's own test/test_iterations.jl and test/test_pulse_optimization.jl) run a second time
WITHOUT a GPU: the product's host driver on the oracle-backed engine of tests/oracle_engine.py.  What this adds to
the `-m "not gpu"` suite is the host logic of optimize.py / workspace.py / result.py; the arithmetic is the oracle's."""
import importlib

import pytest

import oracle_engine as E
import test_parity_gpu as G


@pytest.fixture(autouse=True)
def _oracle_engine(monkeypatch, request):
    monkeypatch.setattr(importlib.import_module("krotov_jl_b200.workspace"), "KrotovCuda", E.OracleEngine)
    before = E.OracleEngine.created
    yield
    if "argument_errors" not in request.node.name:  # (those raise before an engine is made)
        assert E.OracleEngine.created > before  # the stand-in was what ran


test_iter_start_stop_records = G.test_iter_start_stop_records
test_callbacks_order_records_and_pulse_mutation = G.test_callbacks_order_records_and_pulse_mutation
test_pulses_as_controls_are_not_mutated = G.test_pulses_as_controls_are_not_mutated
test_continue_from_and_check_convergence = G.test_continue_from_and_check_convergence
test_atexit_filename_and_pickling_a_result_from_a_callback = G.test_atexit_filename_and_pickling_a_result_from_a_callback
test_continue_from_a_foreign_result = G.test_continue_from_a_foreign_result
test_skip_initial_forward_propagation = G.test_skip_initial_forward_propagation
test_exception_in_callback_is_captured = G.test_exception_in_callback_is_captured
test_storages_are_reachable_from_callbacks = G.test_storages_are_reachable_from_callbacks
test_user_chi_host_path_equals_builtin = G.test_user_chi_host_path_equals_builtin
test_nonuniform_time_grid_weights_and_pulse_options = G.test_nonuniform_time_grid_weights_and_pulse_options
test_trajectory_without_target_uses_host_chi = G.test_trajectory_without_target_uses_host_chi
test_missing_control_derivative_and_two_generators = G.test_missing_control_derivative_and_two_generators
test_c1_tls_parity_and_reference_inequalities = G.test_c1_tls_parity_and_reference_inequalities
test_dense_dummy_problems = G.test_dense_dummy_problems
test_amplitude_argument_errors = G.test_amplitude_argument_errors
test_edge_shapes = G.test_edge_shapes
test_tls_against_the_committed_exact_vectors = G.test_tls_against_the_committed_exact_vectors
test_against_exact_propagator_vectors_of_general_problems = G.test_against_exact_propagator_vectors_of_general_problems
test_seeded_api_variants_vs_oracle = G.test_seeded_api_variants_vs_oracle
test_two_transmon_problem_against_exact_propagator_vector = G.test_two_transmon_problem_against_exact_propagator_vector


def test_nonlinear_amplitudes_through_the_api_against_exact_vector():
    """PolynomialAmplitude / ShapedAmplitude through the product's host path (amplitude collection, `set_amplitudes`)
    against the 40-digit exact-propagator vector of the two-generator problem with a quadratic and a shaped amplitude.
    (CPU engine only: the GPU list of this test was fixed before this vector existed.)"""
    G.test_against_exact_propagator_vectors_of_general_problems("nonlinear_two_generators_d5")
