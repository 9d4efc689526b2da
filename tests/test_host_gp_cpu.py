"""Host logic of the GP / acquisition / BO-loop mirrors under `-m "not gpu"`:
the scenario functions of the GPU suites (ports of gp_test.py, acfun_test.py,
bayesopt_test.py) are re-run here with the engine replaced by
tests/fake_engine.py (the oracle's arithmetic on CPU tensors).  What this
covers is everything ABOVE the C ABI: dataset / cache bookkeeping, noise and
N/(N-1) handling, acquisition callbacks, the BO loops, parameter packing."""
import pytest

from tests import fake_engine
from tests import test_gpu_api as G
from tests import test_quasi_newton as Q


@pytest.mark.parametrize("nx", [20, 0])
def test_predict_identities(monkeypatch, nx):  # gp_test.py:150-207
  fake_engine.install(monkeypatch)
  G.test_predict(nx)


def test_update_dataset_cache_semantics(monkeypatch):  # gp_test.py:209-277
  fake_engine.install(monkeypatch)
  G.test_update_dataset_and_prior_prediction()


@pytest.mark.parametrize("cov", sorted(G.COVS))
def test_kernel_call_signature(monkeypatch, cov):  # kernel_test.py:37-89
  fake_engine.install(monkeypatch)
  G.test_kernel_call_signature(cov)


def test_solve_and_objective_api(monkeypatch):
  fake_engine.install(monkeypatch)
  G.test_solve_gp_linear_system_matches_oracle()


@pytest.mark.parametrize("name", sorted(G.const.ACFUN))
def test_acquisition_shape(monkeypatch, name):  # acfun_test.py:43-72
  fake_engine.install(monkeypatch)
  G.test_acquisition_shape(name)


def test_acquisition_values_and_callbacks(monkeypatch):
  fake_engine.install(monkeypatch)
  G.test_acquisition_values_match_oracle()


def test_bo_loops(monkeypatch):  # bayesopt.py:137-193, bayesopt_test.py:45-103
  fake_engine.install(monkeypatch)
  G.test_simulated_bo_loop_shape()
  Q.test_simulated_bo_picks_the_oracle_argmax()
  for name in ("expected_improvement", "ucb", "random_search"):
    Q.test_run_bayesopt_synthetic(name)


def test_quasi_newton_training(monkeypatch):  # gp.py:158-191
  fake_engine.install(monkeypatch)
  Q.test_infer_parameters_lbfgs_matches_oracle_driven_run(False)
  Q.test_infer_parameters_lbfgs_matches_oracle_driven_run(True)


@pytest.mark.parametrize("cov", sorted(G.COVS))
def test_infer_parameters_decreases_nll(monkeypatch, cov):  # gp_test.py:48-148
  fake_engine.install(monkeypatch)
  G.test_infer_parameters_decreases_nll(cov)


def test_adam_loop_semantics(monkeypatch):  # gp.py:114-157
  fake_engine.install(monkeypatch)
  G.test_objective_api_and_value_and_grad()
  G.test_infer_parameters_matches_oracle_adam_loop()
  G.test_infer_parameters_subsampling_and_nan_at_step0()


def test_objectives_scenarios(monkeypatch):
  """tests/test_gpu_objectives.py on the oracle-backed engine."""
  from tests import test_gpu_objectives as J
  from tests import helpers as H
  fake_engine.install(monkeypatch)
  for name in H.golden_cases(kl=True):
    J.test_divergence_values_match_golden(name)
    J.test_kl_value_and_grad_match_golden(name)
  J.test_nll_plus_regkl_value_and_grad("matern32")
  J.test_scalar_lengthscale_and_no_warp()
  J.test_divergence_edge_cases()
  J.test_training_on_ekl_decreases_it("adam")
  J.test_training_on_ekl_decreases_it("lbfgs")
  J.test_gp_stats()
  J.test_hgp_stats_average_over_samples()
  J.test_sample_mean_cov_regularizer("squared_exponential")


@pytest.mark.parametrize("method", ["adam", "lbfgs"])
def test_train_checkpoint(monkeypatch, tmp_path, method):  # gp.py:151-157,186-191
  fake_engine.install(monkeypatch)
  G.test_train_writes_a_checkpoint(tmp_path, method)
