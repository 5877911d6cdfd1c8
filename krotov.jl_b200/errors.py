"""Exception types named after the Julia exceptions the reference throws, so that tests read alike:
``error("...")`` -> ``ErrorException`` (``src/workspace.jl:72,98``), ``ArgumentError``
(``src/workspace.jl:166-167``, ``src/optimize.jl:163-166``)."""


class ErrorException(RuntimeError):
    pass


class ArgumentError(ValueError):
    pass
