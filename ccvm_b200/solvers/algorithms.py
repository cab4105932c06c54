"""Parameter classes of the specialised in-loop algorithms (reference solvers/algorithms.py)."""


class AdamParameters:
    """Hyper-parameters of the Adam-enhanced loops (Kingma & Ba, arXiv:1412.6980).

    Args:
        alpha (float): step size, >= 0.  Default 0.1.
        beta1 (float): first-moment decay, strictly between 0 and 1.  Default 0.9.
        beta2 (float): second-moment decay in (0, 1]; 1 switches the second moment off.
            Default 0.999.
        add_assign (bool): add the bias-corrected update to the raw gradient instead of
            replacing it.  Default True.
    """

    def __init__(self, alpha=0.1, beta1=0.9, beta2=0.999, add_assign=True):
        if alpha < 0.0:
            raise ValueError(f"AdamAlgorithm: Invalid `alpha` value: {alpha}")
        if not 0 < beta1 < 1:
            raise ValueError(f"AdamAlgorithm: Invalid `beta1` value: {beta1}")
        if not 0 < beta2 <= 1:
            raise ValueError(f"AdamAlgorithm: Invalid `beta2` value: {beta2}")
        self.alpha, self.beta1, self.beta2 = alpha, beta1, beta2
        self.add_assign = bool(add_assign)

    def to_dict(self):
        """The four hyper-parameters as a dict (what ``_solve_adam`` receives)."""
        return {"alpha": self.alpha, "beta1": self.beta1, "beta2": self.beta2, "add_assign": self.add_assign}
