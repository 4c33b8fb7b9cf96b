"""Drop-in stand-in for the `xformers` package as far as Styl3R uses it: `xformers.ops.memory_efficient_attention`
(imported unconditionally at src/model/encoder/backbone/croco/blocks.py:25, called at :126-130 and :192-196).
Registered in sys.modules by `styl3r_b200.compat.install()` so that the unmodified reference blocks run their attention
on the B200 kernels (forward: tcgen05 attention; backward: batched tcgen05 GEMMs, styl3r_b200/attention_bwd.py)."""
from . import ops  # noqa: F401
