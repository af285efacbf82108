"""Stand-ins for the two third-party modules reference scripts import next to `xlb`: `warp` and `jax`.

Reference examples do `import warp as wp` / `import jax.numpy as jnp` for a handful of host-side conveniences
(`wp.synchronize()`, `@wp.func` inlet profiles, `wp.to_jax`, `jnp.sqrt(u[0]**2 + ...)` in post-processing), e.g.
examples/performance/mlups_3d.py:4,81 and examples/cfd/flow_past_sphere_3d.py:14-16,83-97,120-133.  Neither package
is a dependency of this framework.  `install()` registers these light modules under those names ONLY when the real
package is not importable, so that such scripts run unchanged; with real warp / jax installed nothing is shadowed.
No LBM arithmetic lives here.
"""

import importlib.util
import sys


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install():
    if _missing("warp"):
        from xlb_b200.compat import warp_standin

        sys.modules["warp"] = warp_standin
    if _missing("jax"):
        from xlb_b200.compat import jax_standin

        sys.modules["jax"] = jax_standin
        sys.modules["jax.numpy"] = jax_standin.numpy
