"""TEST HARNESS (not product code; SURVEY.md 2 marks the demo state-space models out of scope).  Host mirror of dynamicalsystems/filtering/SsmExamples.scala and of SsmModel.generateSeries (SsmModel.scala:18-54): the two
benchmark state-space models of the reference's UKF / GP-UKF tests and the series sampler that feeds them.  Pure host code
(scalar state); the mapping functions take a MATRIX of points, one per row, like every SsmModel of this package."""
from __future__ import annotations

from typing import Optional

import numpy as np

from gp_algos_b200.gp_predictor import GaussianDistribution
from gp_algos_b200.gp_ukf import SsmModel


def SinusoidalSsm() -> SsmModel:
    """SsmExamples.scala:14-25: z' = sin(z), y = sigmoid(z / 3); latent and observation noise 0.1^2."""
    return SsmModel(transitionFuncImpl=lambda u, pts, t: np.sin(np.atleast_2d(pts)),
                    observationFuncImpl=lambda pts, t: 1.0 / (1.0 + np.exp(-np.atleast_2d(pts) / 3.)),
                    latentNoise=np.array([[0.1 * 0.1]]), obsNoise=np.array([[0.1 * 0.1]]))


def KitagawaSsm() -> SsmModel:
    """SsmExamples.scala:27-42: z' = z/2 + 25 z / (1 + z^2), y = 5 sin(2 z); latent noise 0.01^2, observation noise 0.2^2."""
    def transition(u, pts, t):
        p = np.atleast_2d(pts)
        return (p * 0.5) + ((p * 25.) / (1.0 + p * p))
    return SsmModel(transitionFuncImpl=transition, observationFuncImpl=lambda pts, t: np.sin(np.atleast_2d(pts) * 2.) * 5.,
                    latentNoise=np.array([[0.01 * 0.01]]), obsNoise=np.array([[0.2 * 0.2]]))


def generateSeries(model: SsmModel, length: int, initHiddenState, rng: Optional[np.random.Generator] = None):
    """SsmModel.scala:36-54 -> (hidden d x length, observations d_obs x length).  initHiddenState: a vector (Left) or a
    GaussianDistribution to sample it from (Right).  Each step observes the current state, then samples the next one."""
    rng = rng if rng is not None else np.random.default_rng()
    if isinstance(initHiddenState, GaussianDistribution):
        z = rng.multivariate_normal(initHiddenState.mean, initHiddenState.sigma, method="svd")
    else:
        z = np.asarray(initHiddenState, dtype=np.float64).copy()
    d_obs = model.obsNoise.shape[0]
    hidden = np.zeros((len(z), length))
    obs = np.zeros((d_obs, length))
    for it in range(length):
        y = model.observationFuncImpl(z[None, :], it)[0] + rng.multivariate_normal(np.zeros(d_obs), model.obsNoise, method="svd")
        hidden[:, it] = z
        obs[:, it] = y
        z = model.transitionFuncImpl(None, z[None, :], it)[0] + rng.multivariate_normal(np.zeros(len(z)), model.latentNoise, method="svd")
    return hidden, obs
