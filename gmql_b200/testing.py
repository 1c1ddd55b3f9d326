"""Test / tuning helper: environment variables -> per-context options.

The library itself never reads the environment (include/gmql_cuda.h: gmql_ctx_set_option).  Tests and tuning runs find it
convenient to flip a device path with `GMQL_COVER_PATH=tiled pytest ...` or os.environ inside a test; install_env_hooks()
wraps the operator entry points of the ctypes binding so that, before every call, the GMQL_* variables below are copied into
the options of the context the call is made on.  Product code (gmql_b200.executor) does not depend on this module.
"""
import os

ENV_OPTIONS = {
    "GMQL_COVER_PATH": b"cover.path", "GMQL_COVER_LOGW": b"cover.logw", "GMQL_COVER_WAVES": b"cover.waves",
    "GMQL_COVER_STITCH": b"cover.stitch", "GMQL_COVER_SPECULATE": b"cover.speculate", "GMQL_COVER_KERNEL": b"cover.kernel", "GMQL_COVER_DBG": b"cover.dbg", "GMQL_COVER_BLOCKS": b"cover.blocks",
    "GMQL_MAP_INDEX": b"map.index", "GMQL_MAP_GROUPED": b"map.grouped", "GMQL_MAP_LEAN": b"map.lean", "GMQL_DEFER_CHECK": b"defer_check", "GMQL_COVER_GMAP4": b"cover.gmap4", "GMQL_COVER_BINS": b"cover.bins", "GMQL_COVER_FINISH": b"cover.finish", "GMQL_JOIN_MD": b"join.md", "GMQL_JOIN_MD_MEMO": b"join.md_memo",
}
_OPERATORS = ("gmql_cover", "gmql_cover_shard", "gmql_map", "gmql_difference", "gmql_join")


def install_env_hooks(lib):
    """Idempotent.  Returns lib."""
    if getattr(lib, "_gmql_env_hooks", False):
        return lib
    seen = {}

    def sync(ctx):
        key = ctx.value if hasattr(ctx, "value") else ctx
        vals = tuple(os.environ.get(e) for e in ENV_OPTIONS)
        before = seen.get(key, (None,) * len(vals))
        if before == vals:
            return
        # only what the environment sets (or has set and dropped again) is touched: options the program chose itself through
        # gmql_ctx_set_option (bench.py: defer_check) stay as they are
        for (env, name), v, old in zip(ENV_OPTIONS.items(), vals, before):
            if v != old:
                lib.gmql_ctx_set_option(ctx, name, v.encode() if v else None)
        seen[key] = vals

    def wrap(raw):
        def call(ctx, *args):
            sync(ctx)
            return raw(ctx, *args)
        return call

    for name in _OPERATORS:
        setattr(lib, name, wrap(getattr(lib, name)))
    raw_destroy = lib.gmql_ctx_destroy

    def destroy(ctx):
        seen.pop(ctx.value if hasattr(ctx, "value") else ctx, None)
        return raw_destroy(ctx)
    lib.gmql_ctx_destroy = destroy
    lib._gmql_env_hooks = True
    return lib
