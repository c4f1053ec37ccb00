"""gymnasium.error stand-in (test infrastructure)."""


class Error(Exception):
    pass
