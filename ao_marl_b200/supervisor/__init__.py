from .rlSupervisor import RlSupervisor  # noqa: F401
