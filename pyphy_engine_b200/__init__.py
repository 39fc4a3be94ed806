"""pyphy_engine_b200 -- batched billiards simulate-and-render on B200 (sm_100a).

Reference-facing modules (same names and call shapes as pulkitag/pyphy-engine):
``geometry``, ``primitives``, ``dynamics``, ``dataio``, ``save_data``.
Batched engine: ``engine.WorldBatch`` over the C ABI in ``include/pyphy_b200.h``
(library built by ``python -m pyphy_engine_b200.build``).  Importing the package does not load
CUDA; the first object that needs the device does, and fails loudly if the library is missing.
"""
__version__ = "0.1.0"
