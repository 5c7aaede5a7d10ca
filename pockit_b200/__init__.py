"""pockit-b200: a B200-native evaluation engine for pockit's NLP callbacks.

``pockit_b200.lobatto`` / ``pockit_b200.radau`` expose ``System`` and ``Phase``
with the reference's modelling API (``pockit.lobatto`` / ``pockit.radau``); the
callbacks run on the GPU through the C-ABI library ``libpockit_b200.so``.
"""

__version__ = "0.1.0"
