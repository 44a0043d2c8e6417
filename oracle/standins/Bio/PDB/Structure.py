"""Stand-in submodule (utils/modeler.py:20); only the name is needed to import the reference."""


class Structure:
    def __init__(self, *a, **k):
        raise NotImplementedError("I/O-only stand-in")
