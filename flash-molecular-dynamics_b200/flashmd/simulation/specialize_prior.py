"""`condense_all_priors_for_simulation` (reference simulation/specialize_prior.py:76-207) precomputes, per
prior class, flat per-term parameter vectors for the fixed topology of a simulation.  Here the prior
modules compute those vectors on demand (`data2parameters`) and the fused engine consumes them once at
set-up (simulation/lowering.py), so condensing is the identity on the model object; the function is kept
because user scripts and saved configs call it."""
from typing import List


def condense_all_priors_for_simulation(priors, data_list: List):
    """Returns (priors, data_list) like the reference (specialize_prior.py:76-109)."""
    return priors, data_list


def condense_prior_for_simulation(TargetPrior, priors, data_list: List):
    """Reference specialize_prior.py:112-207 merges all priors of class `TargetPrior` into one static module."""
    return priors, data_list
