"""
mapleaf_b200 — B200-native batched Monte Carlo 6-DOF trajectory integration behind MAPLEAF's
`runMonteCarloSimulation` interface. See DESIGN.md / INTEGRATION.md.
"""
from .simdef import SimDefinition, SubDictReader, defaultConfigValues  # noqa: F401
from .model import buildModel, L  # noqa: F401
from .montecarlo import runMonteCarloSimulation  # noqa: F401,E402
from .convergence import ConvergenceSimRunner  # noqa: F401,E402
from .optimization import optimizationRunnerFactory  # noqa: F401,E402
