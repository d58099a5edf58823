"""Exception types mirroring ``cashocs/_exceptions.py`` (names and meaning)."""

from __future__ import annotations


class CashocsException(Exception):
    """Base exception (``cashocs/_exceptions.py:25``)."""


class NotConvergedError(CashocsException):
    """A solver did not converge (``cashocs/_exceptions.py:33-59``)."""

    def __init__(self, solver: str, message: str | None = None) -> None:
        super().__init__()
        self.solver = solver
        self.message = message

    def __str__(self) -> str:
        tail = f"\n{self.message}" if self.message else ""
        return f"The {self.solver} failed to converge.{tail}"


class PETScError(CashocsException):
    """Base class of errors raised by the (PETSc-compatible) linear solvers."""


class PETScKSPError(PETScError):
    """The Krylov solver ended with a negative converged reason.

    Same reason codes and messages as ``cashocs/_exceptions.py:93-129`` so callers that catch
    the reference exception and inspect ``error_code`` keep working.
    """

    error_dict = {
        -2: " (ksp_diverged_null)",
        -3: ", reached maximum iterations (ksp_diverged_its)",
        -4: ", reached divergence tolerance (ksp_diverged_dtol)",
        -5: ", krylov method breakdown (ksp_diverged_breakdown)",
        -6: " (ksp_diverged_breakdown_bicg)",
        -7: ", require nonsymmetric matrix (ksp_diverged_nonsymmetric)",
        -8: ", the preconditioner is indefinite (ksp_diverged_indefinite_pc)",
        -9: ", a norm of the residual is NaN or Inf (ksp_diverged_nanorinf)",
        -10: ", indefinite matrix (ksp_diverged_indefinite_mat)",
        -11: ", failed to setup preconditioner (ksp_diverged_pc_failed)",
    }

    def __init__(self, error_code: int, message: str | None = None) -> None:
        super().__init__()
        self.error_code = error_code
        self.message = message
        self.error_message = self.error_dict.get(error_code, " (unknown)")

    def __str__(self) -> str:
        tail = f"\n{self.message}" if self.message else ""
        return f"The PETSc linear solver did not converge{self.error_message}.{tail}"


class InputError(CashocsException):
    """Invalid user input (``cashocs/_exceptions.py:132-170``)."""

    def __init__(self, obj: str, param: str, message: str | None = None) -> None:
        super().__init__()
        self.obj = obj
        self.param = param
        self.message = message

    def __str__(self) -> str:
        tail = f"\n{self.message}" if self.message else ""
        return f"Not a valid input for object {self.obj}. The faulty input is for the parameter {self.param}.{tail}"


class ConfigError(CashocsException):
    """Invalid configuration (``cashocs/_exceptions.py:173-201``)."""

    pre_message = "You have some error(s) in your config file.\n"

    def __init__(self, config_errors: list[str]) -> None:
        super().__init__()
        self.config_errors = config_errors

    def __str__(self) -> str:
        return self.pre_message + "".join(self.config_errors)
