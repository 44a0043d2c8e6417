"""Stand-in submodule (utils/modeler.py:21); only the name is needed to import the reference."""


class Model:
    def __init__(self, *a, **k):
        raise NotImplementedError("I/O-only stand-in")
