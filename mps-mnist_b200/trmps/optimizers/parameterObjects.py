"""Knob holders with the reference's names and defaults (trmps/optimizers/parameterObjects.py:1-74)."""


class MPSOptimizerParameters(object):
    """cutoff and armijo_iterations are accepted for compatibility; like the reference, the optimizer
    ignores them (the Armijo bound is 19 trials, baseoptimizer.py:318)."""

    def __init__(self, cutoff=1000, reg=0.001, lr_reg=0.99, min_singular_value=10 ** (-4), verbosity=0,
                 armijo_coeff=10 ** (-4), use_hessian=False, armijo_iterations=10, path="MPSconfig",
                 updates_per_step=1):
        self.cutoff = cutoff
        self.reg = reg
        self.lr_reg = lr_reg
        self.min_singular_value = min_singular_value
        self.verbosity = verbosity
        self.armijo_coeff = armijo_coeff
        self.use_hessian = use_hessian
        self.armijo_iterations = armijo_iterations
        self.path = path
        self.updates_per_step = updates_per_step


class MPSTrainingParameters(object):
    def __init__(self, rate_of_change=1000, initial_weights=None, verbose_save=True, _logging_enabled=False):
        self.rate_of_change = rate_of_change
        self.initial_weights = initial_weights
        self._logging_enabled = _logging_enabled
        self.verbose_save = verbose_save
