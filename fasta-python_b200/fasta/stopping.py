"""Stopping rules for fasta(): same names, signature and truth tables as the reference
``fasta/stopping.py`` (reference lines cited per rule).  Each rule is called once per iteration
as ``stop_rule(i, resid, norm_resid, max_resid, tolerance)`` with ``np.float64`` arguments
(reference ``fasta/__init__.py:308``); division by a zero ``max_resid`` therefore yields
inf/nan with a numpy warning instead of raising, exactly like the reference.
"""

__all__ = ["residual", "norm_residual", "ratio_residual", "hybrid_residual"]


def residual(i, resid, norm_resid, max_resid, tolerance):
    """True once the residual |Dx|/tau drops below the tolerance (reference stopping.py:6-15)."""
    return resid < tolerance


def norm_residual(i, resid, norm_resid, max_resid, tolerance):
    """True once the normalised residual drops below the tolerance (reference stopping.py:18-27)."""
    return norm_resid < tolerance


def ratio_residual(i, resid, norm_resid, max_resid, tolerance):
    """True once residual / largest-residual-so-far drops below the tolerance (reference stopping.py:30-39)."""
    return resid / max_resid < tolerance


def hybrid_residual(i, resid, norm_resid, max_resid, tolerance):
    """ratio_residual OR norm_residual -- the default rule (reference stopping.py:42-51, __init__.py:45)."""
    return resid / max_resid < tolerance or norm_resid < tolerance
