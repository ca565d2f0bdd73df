"""TEST/BENCH INFRASTRUCTURE ONLY — torch-CPU port of the reference's detector forward passes.

This is what ``bench.py`` times for ``cpu_baseline`` and ``--impl reference`` (kind "port"): the
reference is pure Python on torch and cannot travel to the GPU box, so its forward passes are
restated here op for op (same torch operator sequence per trellis stage, so the CPU cost profile —
MKL GEMMs for the MLP, ~8 small torch ops per stage for the ACS — is the reference's own):
  acs stage      python_code/utils/trellis_utils.py:16-30
  VNET forward   python_code/detectors/VNET/vnet_detector.py:35-63
  VA forward     python_code/detectors/VA/va_detector.py:52-98 (tap table passed in)
It is validated against the numpy oracle (and through it against the golden fixtures) in
tests/test_oracle_golden.py::test_torch_port_matches_oracle.  Never imported by the product.
"""
import math

import torch


def acs_stage(pm: torch.Tensor, cost: torch.Tensor, table_flat: torch.Tensor, n_states: int):
    """trellis_utils.py:26-30: build the two index vectors, gather pm+cost, view [B,S,2], min."""
    batch = pm.size(0)
    src = table_flat.repeat(batch).long()
    rows = torch.arange(batch).repeat_interleave(2 * n_states)
    candidates = (pm + cost)[rows, src].reshape(-1, n_states, 2)
    return torch.min(candidates, dim=2)


def transition_table_flat(n_states: int) -> torch.Tensor:
    """trellis_utils.py:12 flattened: [0,1,2,...,S-1,0,1,...,S-1] as float like the detectors keep it."""
    return torch.cat([torch.arange(n_states), torch.arange(n_states)]).float()


def stage_loop(cost: torch.Tensor, n_stages: int) -> torch.Tensor:
    """va_detector.py:89-98 / vnet_detector.py:51-61."""
    batch, width, n_states = cost.shape
    table = transition_table_flat(n_states)
    pm = torch.zeros([batch, n_states])
    out = torch.zeros([batch, width])
    for i in range(n_stages):
        out[:, i] = torch.argmin(pm, dim=1) % 2
        pm, _ = acs_stage(pm, cost[:, i], table, n_states)
    return out


def make_net(n_states: int) -> torch.nn.Sequential:
    """vnet_detector.py:27-33."""
    return torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50),
                               torch.nn.ReLU(), torch.nn.Linear(50, n_states))


def load_weights(net: torch.nn.Sequential, weights) -> None:
    with torch.no_grad():
        for p, w in zip(net.parameters(), weights):
            p.copy_(torch.as_tensor(w).reshape(p.shape))


def vnet_forward_val(net: torch.nn.Sequential, y: torch.Tensor, n_stages: int) -> torch.Tensor:
    """vnet_detector.py:46-61."""
    n_states = net[4].out_features
    priors = net(y.reshape(-1, 1)).reshape(y.shape[0], y.shape[1], n_states)
    return stage_loop(-priors, n_stages)


def va_forward_val(y: torch.Tensor, state_priors_sn: torch.Tensor, n_stages: int) -> torch.Tensor:
    """va_detector.py:62-68,89-98; state_priors_sn is compute_state_priors' [S, n_h] table."""
    reps = y.shape[0] // state_priors_sn.shape[1]
    cost = y.unsqueeze(dim=2) - state_priors_sn.T.repeat(repeats=[reps, 1]).unsqueeze(dim=1)
    cost = cost ** 2 / 2 - math.log(math.sqrt(2 * math.pi))
    return stage_loop(cost, n_stages)
