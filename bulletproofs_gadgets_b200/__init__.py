"""bulletproofs_gadgets_b200 -- B200-native hot path of MarcKloter/bulletproofs_gadgets.

Host-side mirror (Python, over the C ABI of libbpg.so) of the reference-facing interface for the path:
PedersenGens / BulletproofGens / Prover / Verifier / mimc_hash, with the same argument meaning and error
behaviour as the bulletproofs fork the reference links."""
from ._lib import BpgError, load  # noqa: F401
from .api import (BulletproofGens, Context, PedersenGens, Prover, R1CSError, Transcript, Verifier, mimc_hash,  # noqa: F401
                  mimc_hash_batch, merkle_node_batch)
