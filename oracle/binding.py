"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liboracle.so (the CPU restatement of the reference
shaders).  May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
only; the product package never touches it."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
P = C.c_void_p


def oracle_lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `make oracle` (or __graft_entry__.build())")
    lib = C.CDLL(path)
    sig = {
        "orc_set_threads": (C.c_int, [C.c_int]),
        "orc_scene_create": (P, [P]),
        "orc_scene_destroy": (None, [P]),
        "orc_scene_set_brute_force": (None, [P, C.c_int]),
        "orc_scene_set_prev_instances": (None, [P, P, C.c_uint32]),
        "orc_scene_num_triangles": (C.c_uint32, [P]),
        "orc_frame_create": (P, [C.c_uint32, C.c_uint32]),
        "orc_frame_destroy": (None, [P]),
        "orc_frame_clear": (None, [P]),
        "orc_frame_flip": (None, [P]),
        "orc_set_camera": (None, [P, P, P]),
        "orc_gbuffer": (None, [P, P]),
        "orc_di_naive": (None, [P, P]),
        "orc_di_naive_rt": (None, [P, P]),
        "orc_gi_naive": (None, [P, P]),
        "orc_di_pathgen": (None, [P, P, P]),
        "orc_di_temporal": (None, [P, P, P]),
        "orc_di_spatial": (None, [P, P, P]),
        "orc_gi_restir": (None, [P, P]),
        "orc_gris_pathtrace": (None, [P, P, P]),
        "orc_gris_temporal": (None, [P, P, P]),
        "orc_gris_spatial": (None, [P, P, P]),
        "orc_visualize_as": (None, [P, P]),
        "orc_postprocess": (None, [P, P, P]),
        "orc_read": (C.c_int, [P, C.c_int, P, C.c_size_t]),
        "orc_write": (C.c_int, [P, C.c_int, P, C.c_size_t]),
        "orc_trace_closest": (None, [P, P, C.c_uint32, P]),
        "orc_trace_shadow": (None, [P, P, C.c_uint32, P]),
        "orc_counters_reset": (None, [P]),
        "orc_counters_read": (None, [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "orc_hash2": (C.c_uint32, [C.c_uint32]),
        "orc_make_seed": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_sample1f": (C.c_float, [C.POINTER(C.c_uint32)]),
        "orc_sincos": (None, [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "orc_round_through_half": (C.c_float, [C.c_float]),
        "orc_concentric_disk": (None, [C.c_float, C.c_float, C.POINTER(C.c_float)]),
        "orc_sample_light": (None, [P, P, P, P, P, P, P, P, P, P]),
        "orc_is_bsdf_delta": (C.c_int, [P]),
        "orc_is_bsdf_connectible": (C.c_int, [P]),
        "orc_eval_bsdf": (None, [P, P, P, P, P, P, C.POINTER(C.c_float)]),
        "orc_sample_bsdf": (C.c_int, [P, P, P, P, P, P, P, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
