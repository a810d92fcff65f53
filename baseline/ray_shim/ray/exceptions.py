"""Exception types the reference imports from ray.exceptions."""


class RayError(Exception):
    pass


class RayTaskError(RayError):
    pass


class RayActorError(RayError):
    pass


class GetTimeoutError(RayError, TimeoutError):
    pass
