"""ctypes loader for libtmb200.so (the C ABI in include/tmb200.h).

The product path has NO fallback: if the CUDA library has not been built, importing the API fails loudly.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# TMB200_LIB: load another build of the same library (A/B measurements of kernel variants); the product path is the default
LIB_PATH = Path(os.environ.get("TMB200_LIB") or (Path(__file__).resolve().parent / "lib" / "libtmb200.so"))

TMB_OK = 0
TMB_ERROR = 1
TMB_INVALID_HARDWARE = 2
F32 = 32
F64 = 64


class I128(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_int64)]


_p_f64 = C.POINTER(C.c_double)
_p_f32 = C.POINTER(C.c_float)
_p_i32 = C.POINTER(C.c_int32)
_p_u32 = C.POINTER(C.c_uint32)
_p_u64 = C.POINTER(C.c_uint64)
_p_i128 = C.POINTER(I128)
_h = C.c_void_p
_ph = C.POINTER(C.c_void_p)
_int = C.c_int
_dbl = C.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES).  Must stay in sync with include/tmb200.h;
# tests/test_abi.py parses the header and checks every declared symbol is exported and listed here.
SIGNATURES = {
    "tmb_cuda_device_reset": [],
    "tmb_set_stream": [_h],
    "tmb_device_synchronize": [],
    "tmb_harmonic_bond_create": [_int, _p_i32, _int, _ph],
    "tmb_harmonic_angle_create": [_int, _p_i32, _int, _ph],
    "tmb_periodic_torsion_create": [_int, _p_i32, _int, _ph],
    "tmb_flat_bottom_bond_create": [_int, _p_i32, _int, _ph],
    "tmb_centroid_restraint_create": [_int, _p_i32, _int, _p_i32, _int, _dbl, _dbl, _ph],
    "tmb_log_flat_bottom_bond_create": [_int, _p_i32, _int, _dbl, _ph],
    "tmb_chiral_atom_restraint_create": [_int, _p_i32, _int, _ph],
    "tmb_chiral_bond_restraint_create": [_int, _p_i32, _int, _p_i32, _int, _ph],
    "tmb_nonbonded_pair_list_precomputed_create": [_int, _p_i32, _int, _dbl, _dbl, _ph],
    "tmb_nonbonded_all_pairs_create": [_int, _int, _dbl, _dbl, _p_i32, _int, _int, _dbl, _ph],
    "tmb_nonbonded_all_pairs_set_atom_idxs": [_h, _p_i32, _int],
    "tmb_nonbonded_all_pairs_get_num_atom_idxs": [_h, C.POINTER(_int)],
    "tmb_nonbonded_all_pairs_get_atom_idxs": [_h, _p_i32],
    "tmb_nonbonded_interaction_group_create": [_int, _int, _p_i32, _int, _dbl, _dbl, _p_i32, _int, _int, _dbl, _ph],
    "tmb_nonbonded_interaction_group_set_atom_idxs": [_h, _p_i32, _int, _p_i32, _int],
    "tmb_nonbonded_pair_list_create": [_int, _int, _p_i32, _int, _p_f64, _int, _dbl, _dbl, _ph],
    "tmb_summed_potential_create": [_ph, _int, _p_i32, _int, _int, _ph],
    "tmb_fanout_summed_potential_create": [_ph, _int, _int, _ph],
    "tmb_potential_destroy": [_h],
    "tmb_potential_execute": [_h, _int, _int, _p_f64, _p_f64, _p_f64, _p_u64, _p_u64, _p_i128],
    "tmb_potential_execute_batch": [_h, _int, _int, _int, _int, _p_f64, _p_f64, _p_f64, _p_u64, _p_u64, _p_i128],
    "tmb_potential_execute_batch_sparse": [
        _h, _int, _int, _int, _int, _int, _p_u32, _p_u32, _p_f64, _p_f64, _p_f64, _p_u64, _p_u64, _p_i128,
    ],
    "tmb_potential_du_dp_fixed_to_float": [_h, _int, _int, _p_u64, _p_f64],
    "tmb_potential_execute_device": [_h, _int, _int, _h, _h, _h, _h, _h, _h, _h],
    "tmb_nonbonded_num_tiles": [_h, C.POINTER(C.c_uint)],
    "tmb_nonbonded_num_rebuilds": [_h, C.POINTER(C.c_uint)],
    "tmb_nonbonded_tile_capacity": [_h, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)],
    "tmb_nonbonded_set_kernel_timing": [_h, _int],
    "tmb_nonbonded_drain_kernel_times": [_h, _p_f32, _int, C.POINTER(_int)],
    "tmb_bound_potential_create": [_h, _p_f64, _int, _ph],
    "tmb_bound_potential_destroy": [_h],
    "tmb_bound_potential_set_params": [_h, _p_f64, _int],
    "tmb_bound_potential_size": [_h, C.POINTER(_int)],
    "tmb_bound_potential_execute": [_h, _int, _p_f64, _p_f64, _p_u64, _p_i128],
    "tmb_bound_potential_execute_batch": [_h, _int, _int, _p_f64, _p_f64, _p_u64, _p_i128],
    "tmb_bound_potential_set_params_device": [_h, _h, _int, _h],
    "tmb_bound_potential_execute_device": [_h, _int, _h, _h, _h, _h, _h],
    "tmb_langevin_integrator_create": [_p_f64, _int, _dbl, _dbl, _dbl, _int, _ph],
    "tmb_langevin_integrator_destroy": [_h],
    "tmb_langevin_integrator_set_noise": [_h, _p_f32],
    "tmb_langevin_integrator_set_step": [_h, C.c_uint64],
    "tmb_langevin_integrator_get_step": [_h, _p_u64],
    "tmb_rmsd_align": [_p_f64, _p_f64, _int, _p_f64],
    "tmb_velocity_verlet_integrator_create": [_dbl, _p_f64, _int, _ph],
    "tmb_integrator_destroy": [_h],
    "tmb_context_initialize": [_h],
    "tmb_context_finalize": [_h],
    "tmb_context_create": [_p_f64, _p_f64, _p_f64, _int, _h, _ph, _int, _ph],
    "tmb_context_create_with_movers": [_p_f64, _p_f64, _p_f64, _int, _h, _ph, _int, _ph, _int, _ph],
    "tmb_barostat_create": [_int, _dbl, _dbl, _p_i32, _p_i32, _int, _int, _ph, _int, _int, _int, _dbl, _ph],
    "tmb_mover_destroy": [_h],
    "tmb_mover_set_interval": [_h, _int],
    "tmb_mover_get_interval": [_h, C.POINTER(C.c_int)],
    "tmb_mover_set_step": [_h, _int],
    "tmb_mover_move_host": [_h, _int, _p_f64, _p_f64, _p_f64, _p_f64],
    "tmb_barostat_set_volume_scale_factor": [_h, _dbl],
    "tmb_barostat_get_volume_scale_factor": [_h, C.POINTER(C.c_double)],
    "tmb_barostat_set_adaptive_scaling": [_h, _int],
    "tmb_barostat_get_adaptive_scaling": [_h, C.POINTER(C.c_int)],
    "tmb_barostat_set_pressure": [_h, _dbl],
    "tmb_barostat_last_uniforms": [_h, _p_f32],
    "tmb_barostat_counters": [_h, C.POINTER(C.c_int)],
    "tmb_bd_exchange_move_create": [_int, _int, _p_i32, _p_i32, _int, _p_f64, _int, _dbl, _dbl, _dbl, _int, _int, _int, _int, _ph],
    "tmb_tibd_exchange_move_create": [_int, _int, _p_i32, _int, _p_i32, _p_i32, _int, _p_f64, _int, _dbl, _dbl, _dbl, _dbl, _int, _int, _int, _int, _ph],
    "tmb_inner_and_outer_mols": [_int, _p_i32, _int, _int, _p_f64, _p_f64, _p_i32, _p_i32, _int, _dbl, _p_i32],
    "tmb_translations_inside_and_outside_sphere": [_int, _int, _p_f64, _p_f64, _dbl, _int, _p_f64],
    "tmb_bd_exchange_move_num_target_mols": [_h, C.POINTER(C.c_int)],
    "tmb_bd_exchange_move_batch_size": [_h, C.POINTER(C.c_int)],
    "tmb_bd_exchange_move_initial_log_weights": [_h, _int, _p_f64, _p_f64, _p_f64],
    "tmb_bd_exchange_move_incremental_log_weights": [_h, _int, _p_f64, _p_f64, _p_i32, _p_f64, _p_f64, _p_f64],
    "tmb_bd_exchange_move_get_params": [_h, _p_f64, _int],
    "tmb_bd_exchange_move_set_params": [_h, _p_f64, _int],
    "tmb_bd_exchange_move_last_log_probability": [_h, C.POINTER(C.c_double)],
    "tmb_bd_exchange_move_last_raw_log_probability": [_h, C.POINTER(C.c_double)],
    "tmb_bd_exchange_move_n_accepted": [_h, C.POINTER(C.c_ulonglong)],
    "tmb_bd_exchange_move_n_proposed": [_h, C.POINTER(C.c_ulonglong)],
    "tmb_bd_exchange_move_before_log_weights": [_h, _p_f64],
    "tmb_bd_exchange_move_after_log_weights": [_h, _p_f64],
    "tmb_nonbonded_mol_energies": [_int, _int, _p_i32, _p_i32, _int, _dbl, _dbl, _p_f64, _p_f64, _p_f64, _p_i128],
    "tmb_atom_by_atom_energies": [_int, _int, _p_i32, _int, _p_f64, _p_f64, _p_f64, _dbl, _dbl, _p_f64],
    "tmb_segmented_logsumexp": [_int, _int, _int, _p_f64, _p_i32, _int, _p_f64],
    "tmb_weighted_sampler_create": [_int, _int, _int, _int, _ph],
    "tmb_weighted_sampler_destroy": [_h],
    "tmb_weighted_sampler_sample": [_h, _p_f64, _p_i32, _int, _p_i32],
    "tmb_rotate_coords": [_int, _int, _int, _p_f64, _p_f64, _p_f64],
    "tmb_rotate_and_translate_mol": [_int, _int, _int, _p_f64, _p_f64, _p_f64, _p_f64, _p_f64],
    "tmb_context_destroy": [_h],
    "tmb_context_step": [_h],
    "tmb_context_multiple_steps": [_h, _int, _int, _p_f64, _p_f64],
    "tmb_context_setup_local_md": [_h, _dbl, _int],
    "tmb_context_multiple_steps_local": [_h, _int, _p_i32, _int, _int, _dbl, _dbl, _int, _p_f64, _p_f64],
    "tmb_context_multiple_steps_local_selection": [_h, _int, _int, _p_i32, _int, _int, _dbl, _dbl, _p_f64, _p_f64],
    "tmb_context_local_md_free_idxs": [_h, _p_u32],
    "tmb_context_set_x_t": [_h, _p_f64],
    "tmb_context_set_v_t": [_h, _p_f64],
    "tmb_context_set_box": [_h, _p_f64],
    "tmb_context_get_x_t": [_h, _p_f64],
    "tmb_context_get_v_t": [_h, _p_f64],
    "tmb_context_get_box": [_h, _p_f64],
    "tmb_context_num_atoms": [_h, C.POINTER(_int)],
    "tmb_context_set_stream": [_h, _h],
    "tmb_context_set_use_graphs": [_h, _int],
    "tmb_context_device_state": [_h, _ph, _ph, _ph],
    "tmb_neighborlist_create": [_int, _int, _ph],
    "tmb_neighborlist_destroy": [_h],
    "tmb_neighborlist_build": [_h, _int, _p_f64, _p_f64, _dbl, C.POINTER(_int), C.POINTER(_int)],
    "tmb_neighborlist_fetch": [_h, _p_i32, _p_i32],
    "tmb_neighborlist_compute_block_bounds": [_h, _int, _p_f64, _p_f64, _p_f64, _p_f64],
    "tmb_neighborlist_set_row_idxs": [_h, _p_u32, _int],
    "tmb_neighborlist_reset_row_idxs": [_h],
    "tmb_neighborlist_resize": [_h, _int],
    "tmb_neighborlist_get_tile_ixn_count": [_h, C.POINTER(C.c_uint)],
    "tmb_neighborlist_get_max_ixn_count": [_h, C.POINTER(_int)],
    "tmb_neighborlist_get_num_row_idxs": [_h, C.POINTER(_int)],
    "tmb_hilbert_sort_create": [_int, _ph],
    "tmb_hilbert_sort_destroy": [_h],
    "tmb_hilbert_sort_sort": [_h, _int, _p_f64, _p_f64, _p_u32],
    "tmb_fill_normal": [_p_f32, _int, C.c_uint64, C.c_uint64],
    "tmb_hrex_run_neighbor_swaps": [_int, _int, _int, _p_i32, _p_f64, _int, _p_i32, _p_f64, _p_i32, _p_u32, _p_u32],
    # non-status functions
    "tmb_last_error": [],
    "tmb_version": [],
    "tmb_fixed_exponent": [],
    "tmb_kernel_launch_count": [],
}

_RESTYPES = {
    "tmb_last_error": C.c_char_p,
    "tmb_version": C.c_int,
    "tmb_fixed_exponent": C.c_uint64,
    "tmb_kernel_launch_count": C.c_longlong,
}

_lib = None


def load() -> C.CDLL:
    """Load libtmb200.so, or raise: there is no CPU / PyTorch fallback for the product path."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -m timemachine_b200.build, or __graft_entry__.build()). There is no fallback path."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib
