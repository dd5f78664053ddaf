"""ExplicitIntegration/Dissipation call surface."""

__all__ = ["artificialDissipationGLF"]


def artificialDissipationGLF(t, data, derivL, derivR, schemeData):
    """Global Lax-Friedrichs dissipation -- ExplicitIntegration/Dissipation/artificial_diss_glf.py:7-111.

    In this library GLF is not a separate pass: the stage kernel forms 0.5*(R-L)*alpha_d, the derivative min/max
    and max alpha_d in registers while the derivatives are still there (that is the point of the fusion), so this
    callable is the *token* ``schemeData.dissFunc`` must hold.  ``termLaxFriedrichs`` returns the same
    ``stepBound`` the reference's call would."""
    raise NotImplementedError(
        "artificialDissipationGLF is fused into the stage kernel; call termLaxFriedrichs / odeCFL3 "
        "(no standalone CPU evaluation exists in this library)")
