"""Exceptions of the result getters (the reference's names, time_evolution.py:585-595)."""


class ThetaNotStored(Exception):
    pass


class CurrentNotStored(Exception):
    pass


class VoltageNotStored(Exception):
    pass


class DataAtTimepointNotStored(Exception):
    pass
