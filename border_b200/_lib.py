"""ctypes binding of include/border_b200.h (the C-ABI drop-in boundary).

The product path FAILS LOUDLY when the CUDA library is missing: there is no CPU fallback and
nothing here imports or calls oracle/.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libborder_b200.so")

# enums (include/border_b200.h)
BB_U8, BB_F32, BB_I64, BB_I32 = 0, 1, 2, 3
BB_NORM_ALL, BB_NORM_BATCH = 0, 1
BB_LOSS_MSE, BB_LOSS_SMOOTH_L1 = 0, 1
BB_OPT_ADAM, BB_OPT_ADAMW = 0, 1
BB_EXPLORER_SOFTMAX, BB_EXPLORER_EPS_GREEDY = 0, 1
BB_NET_MLP, BB_NET_ATARI_CNN = 0, 1
BB_ENTCOEF_FIX, BB_ENTCOEF_AUTO = 0, 1
(BB_IQN_CONST10, BB_IQN_UNIFORM8, BB_IQN_UNIFORM10, BB_IQN_UNIFORM32, BB_IQN_UNIFORM64, BB_IQN_MEDIAN,
 BB_IQN_CONST32) = range(7)


class bb_replay_cfg(C.Structure):
    _fields_ = [("capacity", C.c_uint64), ("seed", C.c_uint64), ("per_config_some", C.c_int32),
                ("alpha", C.c_float), ("beta_0", C.c_float), ("beta_final", C.c_float),
                ("n_opts_final", C.c_uint64), ("normalize", C.c_int32), ("obs_kind", C.c_int32),
                ("obs_elems", C.c_uint32), ("act_kind", C.c_int32), ("act_elems", C.c_uint32),
                ("fastrand_seed", C.c_uint64), ("device", C.c_int32)]


class bb_batch_view(C.Structure):
    _fields_ = [("batch_size", C.c_uint64), ("obs", C.c_void_p), ("act", C.c_void_p), ("next_obs", C.c_void_p),
                ("reward", C.c_void_p), ("is_terminated", C.c_void_p), ("is_truncated", C.c_void_p),
                ("ix_sample", C.c_void_p), ("weight", C.c_void_p)]


class bb_net_cfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in_dim", C.c_int32), ("n_units", C.c_int32), ("units", C.c_int32 * 8),
                ("out_dim", C.c_int32), ("activation_out", C.c_int32), ("n_stack", C.c_int32),
                ("skip_linear", C.c_int32)]


class bb_opt_cfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("wd", C.c_double), ("eps", C.c_double), ("amsgrad", C.c_int32)]


class bb_dqn_cfg(C.Structure):
    _fields_ = [("q_config", bb_net_cfg), ("opt_config", bb_opt_cfg), ("soft_update_interval", C.c_uint64),
                ("n_updates_per_opt", C.c_uint64), ("batch_size", C.c_uint64), ("discount_factor", C.c_double),
                ("tau", C.c_double), ("train", C.c_int32), ("explorer", C.c_int32), ("eps_start", C.c_double),
                ("eps_final", C.c_double), ("final_step", C.c_uint64), ("clip_reward_some", C.c_int32),
                ("clip_reward", C.c_double), ("double_dqn", C.c_int32), ("clip_td_err_some", C.c_int32),
                ("clip_td_err_min", C.c_double), ("clip_td_err_max", C.c_double), ("device", C.c_int32),
                ("critic_loss", C.c_int32), ("record_verbose_level", C.c_uint64), ("init_seed", C.c_uint64),
                ("explorer_seed", C.c_uint64)]


class bb_sac_cfg(C.Structure):
    _fields_ = [("pi_config", bb_net_cfg), ("pi_opt_config", bb_opt_cfg), ("q_config", bb_net_cfg),
                ("q_opt_config", bb_opt_cfg), ("gamma", C.c_double), ("tau", C.c_double),
                ("ent_coef_mode", C.c_int32), ("ent_coef_fix", C.c_double), ("ent_coef_target", C.c_double),
                ("ent_coef_lr", C.c_double), ("epsilon", C.c_double), ("min_lstd", C.c_double),
                ("max_lstd", C.c_double), ("n_updates_per_opt", C.c_uint64), ("batch_size", C.c_uint64),
                ("train", C.c_int32), ("critic_loss", C.c_int32), ("reward_scale", C.c_double),
                ("n_critics", C.c_uint64), ("seed_some", C.c_int32), ("seed", C.c_int64), ("device", C.c_int32),
                ("init_seed", C.c_uint64), ("noise_seed", C.c_uint64)]


class bb_iqn_cfg(C.Structure):
    _fields_ = [("f_config", bb_net_cfg), ("m_config", bb_net_cfg), ("opt_config", bb_opt_cfg),
                ("feature_dim", C.c_int32), ("embed_dim", C.c_int32), ("soft_update_interval", C.c_uint64),
                ("n_updates_per_opt", C.c_uint64), ("batch_size", C.c_uint64), ("discount_factor", C.c_double),
                ("tau", C.c_double), ("train", C.c_int32), ("sample_percents_pred", C.c_int32),
                ("sample_percents_tgt", C.c_int32), ("sample_percents_act", C.c_int32), ("eps_start", C.c_double),
                ("eps_final", C.c_double), ("final_step", C.c_uint64), ("device", C.c_int32),
                ("init_seed", C.c_uint64), ("explorer_seed", C.c_uint64), ("tau_seed", C.c_uint64)]


class bb_record(C.Structure):
    _fields_ = [("loss", C.c_float), ("loss_critic", C.c_float), ("loss_actor", C.c_float),
                ("ent_coef", C.c_float), ("pred_mean", C.c_float), ("tgt_mean", C.c_float),
                ("reward_mean", C.c_float), ("tgt_minus_pred_mean", C.c_float), ("n_opts", C.c_uint64)]


# every symbol include/border_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SIGS = {
    "bb_last_error": (C.c_char_p, []),
    "bb_abi_version": (C.c_int32, []),
    "bb_device_error": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32]),
    "bb_device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
    "bb_test_powf": (C.c_int32, [C.c_int32, _P, _P, _P, C.c_size_t]),
    "bb_test_gemm": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P,
                                 C.c_int32, _P]),
    "bb_test_conv": (C.c_int32, [C.c_int32] * 10 + [_P] * 5),
    "bb_tma_trace": (C.c_int32, [_P]),
    "bb_tma_trace_ctas": (C.c_int32, [_P]),
    "bb_tma_stats": (C.c_int32, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int32]),
    "bb_bench_gemm": (C.c_int32, [C.c_int32] * 7 + [C.POINTER(C.c_float)]),
    "bb_debug_tc_trace": (C.c_int32, [_P]),
    "bb_bench_conv1": (C.c_int32, [C.c_int32] * 4 + [C.POINTER(C.c_float)]),
    "bb_replay_cfg_default": (None, [C.POINTER(bb_replay_cfg)]),
    "bb_replay_create": (C.c_int32, [C.POINTER(bb_replay_cfg), C.POINTER(_P)]),
    "bb_replay_destroy": (C.c_int32, [_P]),
    "bb_replay_set_stream": (C.c_int32, [_P, _P]),
    "bb_replay_push": (C.c_int32, [_P, _P, _P, _P, _P, _P, _P, C.c_size_t, C.c_int32]),
    "bb_replay_len": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "bb_replay_sample": (C.c_int32, [_P, C.c_size_t, C.POINTER(bb_batch_view)]),
    "bb_replay_batch_to_host": (C.c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "bb_replay_last_batch": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "bb_replay_update_priority": (C.c_int32, [_P, _P, _P, C.c_size_t, C.c_int32]),
    "bb_replay_inject_uniforms": (C.c_int32, [_P, _P, C.c_size_t]),
    "bb_replay_dump_sum_tree": (C.c_int32, [_P, _P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bb_replay_state": (C.c_int32, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bb_replay_fill_synthetic": (C.c_int32, [_P, C.c_uint64, C.c_uint32, C.c_uint64]),
    "bb_dqn_cfg_default": (None, [C.POINTER(bb_dqn_cfg)]),
    "bb_sac_cfg_default": (None, [C.POINTER(bb_sac_cfg)]),
    "bb_iqn_cfg_default": (None, [C.POINTER(bb_iqn_cfg)]),
    "bb_dqn_create": (C.c_int32, [C.POINTER(bb_dqn_cfg), C.POINTER(_P)]),
    "bb_sac_create": (C.c_int32, [C.POINTER(bb_sac_cfg), C.POINTER(_P)]),
    "bb_iqn_create": (C.c_int32, [C.POINTER(bb_iqn_cfg), C.POINTER(_P)]),
    "bb_agent_destroy": (C.c_int32, [_P]),
    "bb_agent_set_stream": (C.c_int32, [_P, _P]),
    "bb_agent_set_precision": (C.c_int32, [_P, C.c_int32]),
    "bb_agent_set_train": (C.c_int32, [_P, C.c_int32]),
    "bb_agent_is_train": (C.c_int32, [_P, C.POINTER(C.c_int32)]),
    "bb_agent_sample": (C.c_int32, [_P, _P, C.c_size_t, _P]),
    "bb_agent_opt": (C.c_int32, [_P, _P, C.POINTER(bb_record)]),
    "bb_agent_n_opts": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "bb_agent_opt_profiled": (C.c_int32, [_P, _P, C.c_char_p, C.c_size_t]),
    "bb_agent_save_params": (C.c_int32, [_P, C.c_char_p]),
    "bb_agent_load_params": (C.c_int32, [_P, C.c_char_p]),
    "bb_agent_param_count": (C.c_int32, [_P, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bb_agent_param_info": (C.c_int32, [_P, C.c_char_p, C.c_uint64, C.c_char_p, C.c_size_t,
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "bb_agent_get_param": (C.c_int32, [_P, C.c_char_p, C.c_char_p, _P, C.c_size_t]),
    "bb_agent_set_param": (C.c_int32, [_P, C.c_char_p, C.c_char_p, _P, C.c_size_t]),
    "bb_agent_reset_opt_state": (C.c_int32, [_P]),
    "bb_agent_get_opt_state": (C.c_int32, [_P, C.c_char_p, C.c_char_p, _P, _P, C.c_size_t, C.POINTER(C.c_uint64)]),
    "bb_agent_model_info_size": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "bb_agent_model_info": (C.c_int32, [_P, _P, C.c_size_t, C.POINTER(C.c_uint64)]),
    "bb_agent_sync_model": (C.c_int32, [_P, _P, C.c_size_t]),
    "bb_agent_sync_model_from": (C.c_int32, [_P, _P]),
    "bb_agent_inject_noise": (C.c_int32, [_P, C.c_int32, _P, C.c_size_t]),
    "bb_agent_grad_buffer": (C.c_int32, [_P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "bb_agent_exchange_trace": (C.c_int32, [_P, C.POINTER(C.c_uint32)]),
    "bb_actor_step": (C.c_int32, [_P, _P, _P, _P, C.c_float, C.c_int8, C.c_int8, C.POINTER(C.c_int64)]),
    "bb_actor_reset": (C.c_int32, [_P]),
    "bb_actor_step_n": (C.c_int32, [_P, _P, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int32]),
    "bb_actor_step_dev": (C.c_int32, [_P, _P, _P, _P, C.c_float, C.c_int8, C.c_int8, C.POINTER(C.c_int64)]),
    "bb_atari_create": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "bb_atari_destroy": (C.c_int32, [_P]),
    "bb_atari_set_stream": (C.c_int32, [_P, _P]),
    "bb_atari_reset": (C.c_int32, [_P, _P]),
    "bb_atari_step": (C.c_int32, [_P, _P, _P, C.c_float, C.POINTER(C.c_float)]),
    "bb_atari_obs_device": (C.c_int32, [_P, _P, C.POINTER(_P)]),
    "bb_atari_obs_host": (C.c_int32, [_P, _P]),
    "bb_agent_ipc_export": (C.c_int32, [_P, _P, _P]),
    "bb_agent_ipc_connect": (C.c_int32, [_P, C.c_int32, C.c_int32, _P, _P]),
    "bb_kernel_launch_count": (C.c_int32, [C.POINTER(C.c_uint64), C.c_int32]),
}

_lib = None


class BorderB200Error(RuntimeError):
    pass


def lib():
    """Loads libborder_b200.so (built in-tree by `make` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BorderB200Error(
                "%s is missing: build it with `make` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != 0:
        raise BorderB200Error(lib().bb_last_error().decode("utf-8", "replace"))


def declared_symbols():
    return sorted(_SIGS)
