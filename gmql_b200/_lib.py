"""ctypes binding of libgmqlcuda.so (include/gmql_cuda.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgmqlcuda.so")


class GmqlCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gmql_cuda error %d: %s" % (code, msg))
        self.code = code


class Regions(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("location", C.c_int32), ("coord_bits", C.c_int32),
        ("chrom", C.c_void_p), ("start", C.c_void_p), ("stop", C.c_void_p), ("strand", C.c_void_p),
        ("sample", C.c_void_p), ("n_chrom", C.c_int32), ("n_samples", C.c_int32), ("sample_ids", C.c_void_p),
        ("n_cols", C.c_int32), ("cols", C.POINTER(C.c_void_p)), ("valid", C.POINTER(C.c_void_p)),
        ("dup_of", C.c_void_p), ("chrom_bits", C.c_int32), ("sample_offsets", C.c_void_p), ("chrom_span_hint", C.c_void_p), ("stop_is_length", C.c_int32), ("sorted_by_position", C.c_int32),
    ]


class Agg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("col", C.c_int32)]


class Pair(C.Structure):
    _fields_ = [("left_sample", C.c_int32), ("right_sample", C.c_int32)]


class TextSchema(C.Structure):
    _fields_ = [("chr_pos", C.c_int32), ("start_pos", C.c_int32), ("stop_pos", C.c_int32), ("strand_pos", C.c_int32),
                ("n_values", C.c_int32), ("value_pos", C.c_void_p), ("start_minus_one", C.c_int32)]


class TextOutSchema(C.Structure):
    _fields_ = [("one_based", C.c_int32), ("n_values", C.c_int32), ("value_cols", C.c_void_p)]


class CoverShard(C.Structure):
    _fields_ = [("chrom_span", C.c_void_p), ("first_tile", C.c_int64), ("end_tile", C.c_int64)]


class Atom(C.Structure):
    _fields_ = [("kind", C.c_int32), ("arg", C.c_int64)]


# every symbol include/gmql_cuda.h declares (tests check the library exports all of them)
EXPORTS = [
    "gmql_abi_version", "gmql_ctx_create", "gmql_ctx_destroy", "gmql_last_error", "gmql_ctx_sync",
    "gmql_ctx_launch_count", "gmql_host_alloc", "gmql_host_free", "gmql_map", "gmql_cover",
    "gmql_cover_count_samples", "gmql_join", "gmql_result_rows", "gmql_result_column_len",
    "gmql_result_column_elem_size", "gmql_result_column_device", "gmql_result_column_copy",
    "gmql_result_as_regions", "gmql_result_free", "gmql_last_device_ms", "gmql_last_h2d_ms",
    "gmql_dataset_upload", "gmql_dataset_free", "gmql_ctx_set_profiling", "gmql_ctx_profile_report",
    "gmql_difference", "gmql_profile", "gmql_parse_text", "gmql_cover_shard", "gmql_ctx_set_option", "gmql_format_text", "gmql_group",
]

_lib = None


def load_library(path=None):
    """Loads the CUDA back end.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise GmqlCudaError(-5, "libgmqlcuda.so not found at %s — build it with `python -c 'import __graft_entry__ as g; g.build()'`" % p)
    lib = C.CDLL(p)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.gmql_abi_version.restype = C.c_int
    lib.gmql_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.gmql_ctx_destroy.argtypes = [vp]
    lib.gmql_ctx_destroy.restype = None
    lib.gmql_last_error.argtypes = [vp]
    lib.gmql_last_error.restype = C.c_char_p
    lib.gmql_ctx_sync.argtypes = [vp]
    lib.gmql_ctx_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
    lib.gmql_ctx_launch_count.argtypes = [vp]
    lib.gmql_ctx_launch_count.restype = i64
    lib.gmql_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.gmql_host_free.argtypes = [vp]
    lib.gmql_host_free.restype = None
    lib.gmql_map.argtypes = [vp, C.POINTER(Regions), C.POINTER(Regions), C.POINTER(Pair), i64, C.POINTER(Agg), i32, C.POINTER(vp)]
    lib.gmql_difference.argtypes = [vp, C.POINTER(Regions), C.POINTER(Regions), C.POINTER(Pair), i64, i32, C.POINTER(vp)]
    lib.gmql_parse_text.argtypes = [vp, C.c_char_p, i64, i32, C.POINTER(TextSchema), C.POINTER(C.c_char_p), i32, C.POINTER(vp)]
    lib.gmql_format_text.argtypes = [vp, C.POINTER(Regions), C.POINTER(TextOutSchema), C.POINTER(C.c_char_p), C.POINTER(vp)]
    lib.gmql_profile.argtypes = [vp, C.POINTER(Regions), C.POINTER(vp)]
    lib.gmql_group.argtypes = [vp, C.POINTER(Regions), vp, C.POINTER(Agg), i32, C.POINTER(vp)]
    lib.gmql_cover.argtypes = [vp, C.POINTER(Regions), vp, i32, i32, vp, vp, i64, C.POINTER(Agg), i32, C.POINTER(vp)]
    lib.gmql_cover_shard.argtypes = [vp, C.POINTER(Regions), i32, i32, i32, i64, C.POINTER(CoverShard), C.POINTER(vp)]
    lib.gmql_cover_count_samples.argtypes = [vp, C.POINTER(Regions), C.POINTER(i32)]
    lib.gmql_join.argtypes = [vp, C.POINTER(Regions), C.POINTER(Regions), C.POINTER(Pair), i64, C.POINTER(Atom), i32, i32, i64, i64, vp, vp, C.POINTER(vp)]
    lib.gmql_result_rows.argtypes = [vp]
    lib.gmql_result_rows.restype = i64
    lib.gmql_result_column_len.argtypes = [vp, i32]
    lib.gmql_result_column_len.restype = i64
    lib.gmql_result_column_elem_size.argtypes = [vp, i32]
    lib.gmql_result_column_elem_size.restype = i32
    lib.gmql_result_column_device.argtypes = [vp, i32]
    lib.gmql_result_column_device.restype = vp
    lib.gmql_result_column_copy.argtypes = [vp, vp, i32, vp]
    lib.gmql_result_as_regions.argtypes = [vp, vp, C.POINTER(Regions)]
    lib.gmql_result_free.argtypes = [vp, vp]
    lib.gmql_result_free.restype = None
    lib.gmql_last_device_ms.argtypes = [vp]
    lib.gmql_last_device_ms.restype = C.c_double
    lib.gmql_last_h2d_ms.argtypes = [vp]
    lib.gmql_last_h2d_ms.restype = C.c_double
    lib.gmql_dataset_upload.argtypes = [vp, C.POINTER(Regions), C.POINTER(vp), C.POINTER(Regions)]
    lib.gmql_dataset_free.argtypes = [vp, vp]
    lib.gmql_dataset_free.restype = None
    lib.gmql_ctx_set_profiling.argtypes = [vp, C.c_int]
    lib.gmql_ctx_profile_report.argtypes = [vp, C.c_char_p, C.c_size_t]
    if path is None:
        _lib = lib
    return lib


# column ids (include/gmql_cuda.h)
COL_MAP_ROW_OFFSETS, COL_MAP_REF_ROW, COL_MAP_REF_SAMPLE_OFFSETS, COL_MAP_COUNT = 1, 2, 3, 4
COL_BAG_OFFSETS, COL_BAG_ROWS = 5, 6
COL_COV_GROUP, COL_COV_CHROM, COL_COV_START, COL_COV_STOP, COL_COV_STRAND, COL_COV_ACC, COL_COV_J1, COL_COV_J2 = 16, 17, 18, 19, 20, 21, 22, 23
COL_COV_EDGE, COL_COV_RAW_COUNT, COL_COV_RAW_START, COL_COV_RAW_STOP, COL_COV_RAW_SMIN, COL_COV_RAW_EMAX, COL_COV_RAW_SMAX, COL_COV_RAW_EMIN = 24, 25, 26, 27, 28, 29, 30, 31
COL_JOIN_LEFT_ROW, COL_JOIN_RIGHT_ROW, COL_JOIN_PAIR, COL_JOIN_START, COL_JOIN_STOP, COL_JOIN_STRAND, COL_JOIN_CHROM = 32, 33, 34, 35, 36, 37, 38
COL_DIFF_REF_ROW = 48
COL_PROF_LEFTMOST, COL_PROF_RIGHTMOST, COL_PROF_MIN_LENGTH, COL_PROF_MAX_LENGTH, COL_PROF_SUM_LENGTH, COL_PROF_SUM_LENGTH_SQ, COL_PROF_COUNT = 64, 65, 66, 67, 68, 69, 70
COL_TXT_LINE_OFFSET, COL_TXT_STATUS, COL_TXT_CHROM, COL_TXT_START, COL_TXT_STOP, COL_TXT_STRAND = 80, 81, 82, 83, 84, 85
COL_TXT_BYTES = 86
COL_GROUP_ROW, COL_GROUP_COUNT = 96, 97
COL_AGG_VALUE, COL_AGG_VALID = 256, 512
