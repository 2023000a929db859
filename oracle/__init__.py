"""CPU oracle for the batched voice-conversion forward path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, on the CPU, the arithmetic of the reference (achyun/Autoformer)
for the one path this repository accelerates:

* ``oracle.autovc``  -- ``factory/AutoVC.py:18-211`` (Encoder/Decoder/Postnet/AutoVC)
* ``oracle.lstmdv``  -- ``factory/LstmDV.py:4-24``
* ``oracle.melgan``  -- ``melgan/modules.py:72-130`` (ResnetBlock/Generator)
* ``oracle.meta``    -- ``factory/MetaPool.py``, ``factory/MetaConv.py``, ``factory/MLPMixer.py``

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed *CPU* baseline) -- never on the product path.  The product path in
``autoformer_b200`` fails loudly when its CUDA library is missing; it never
falls back to this code.

Pinning
-------
The reference has no tests and no golden vectors of its own,
so the oracle is pinned against *outputs of the reference itself*: the script
``oracle/make_golden.py`` imports the unmodified reference classes from
``/root/reference`` (possible only in the build container), feeds them seeded
weights (``oracle.seeded``) and seeded synthetic inputs, and stores the outputs
under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every oracle
function against those fixtures, so parity is pinned to the reference's own
forward, executed by torch 2.11 CPU fp32.

The arithmetic library is PyTorch's CPU ATen, exactly as in the reference
(which delegates all math to ``torch.nn``); every LSTM additionally has an
explicit gate-by-gate restatement (``lstm_explicit``) that is checked against
the ATen call, so the gate order / bias convention is spelled out in the open.
"""

from .metrics import rel_l2, centred_rel_l2  # noqa: F401
