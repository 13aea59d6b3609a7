"""Neutral workload generators shared by bench.py's two arms (ours / reference) - neither product nor oracle.

The reference arm of bench.py must not import the product package (it would map libpepflow_b200.so into the process),
so the synthetic complexes, the deterministic weights and the reference's state_dict key list are restated here without
any dependency on pepflowww_b200/.  tests/test_host_logic.py checks that these generators and the product's own
(pepflowww_b200.pep_dataloader / pepflowww_b200.utils) produce identical tensors.
"""
from .synthetic import (deterministic_state_dict, reference_state_dict, state_dict_spec, synthetic_batch,  # noqa: F401
                        torsions_mask)
