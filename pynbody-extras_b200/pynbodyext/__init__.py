"""pynbodyext — B200-native drop-in for the *gravity hot path* of pynbody-extras.

Only ``pynbodyext.gravity`` (and the ``pynbodyext._rust`` boundary it calls) exists here; the
calculator framework, filters, transforms, properties, profiles and dask chunks of the
reference are out of scope.
"""
__version__ = "0.1.0+b200"
