"""tnac4o_b200 -- B200 (sm_100a) native implementation of the tnac4o contraction hot path.

Same public names as the reference package (/root/reference/tnac4o/__init__.py:1-2):
``tnac4o``, ``load``, ``load_Jij``, ``round_Jij``, ``minus_Jij``, ``Jij_f2p``, ``energy_Jij``, ``energy_RMF``.
Importing the package loads tnac4o_b200/lib/libtnac4o_b200.so and fails loudly if it has not been built.
"""
from .solver import tnac4o, load  # noqa: F401
from .auxx import load_Jij, round_Jij, minus_Jij, Jij_f2p, energy_Jij, energy_RMF  # noqa: F401
from . import mps  # noqa: F401

__version__ = '0.1.0'
