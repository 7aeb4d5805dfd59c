"""ctypes prototypes of include/kcnn_capi.h (component / nnet level C ABI)."""
import ctypes

c_int, c_float, c_void_p, c_size_t, c_char_p = (ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                                 ctypes.c_size_t, ctypes.c_char_p)
P, I, F, H = c_void_p, c_int, c_float, c_void_p
PP = ctypes.POINTER(c_void_p)
PI = ctypes.POINTER(c_int)
PS = ctypes.POINTER(c_size_t)
MAT = [P, I, I, I]          # pointer, rows, cols, stride

PROTOS = {
    "kcnn_select_gpu": ([c_char_p], c_int),
    "kcnn_set_compute_stream": ([P], None),
    "kcnn_set_math_mode": ([I], None),
    "kcnn_get_math_mode": ([], c_int),
    "kcnn_set_rand_seed": ([ctypes.c_ulonglong], None),
    "kcnn_last_error": ([], c_char_p),
    "kcnn_enable_profile": ([I], None),
    "kcnn_print_profile": ([], None),
    "kcnn_device_bytes_allocated": ([], c_size_t),
    "kcnn_free": ([P], None),
    "kcnn_mat_conv2d": (MAT + MAT + [I] * 6 + MAT + [I], c_int),
    "kcnn_mat_add_mat_rep_vec": (MAT + [P, I, I], c_int),
    "kcnn_mat_flip_mat": (MAT + [I] * 4 + MAT, c_int),
    "kcnn_mat_padding_zero": (MAT + [I] * 5 + MAT, c_int),
    "kcnn_mat_tp_block": (MAT + [I, I] + MAT, c_int),
    "kcnn_mat_tp_inside_block": (MAT + [I, I] + MAT, c_int),
    "kcnn_mat_mod_permute_row": (MAT + [I, I] + MAT, c_int),
    "kcnn_mat_maxpool_prop": (MAT + [I] * 7 + MAT, c_int),
    "kcnn_mat_maxpool_backprop": (MAT + MAT + MAT + MAT + [I] * 7, c_int),
    "kcnn_component_new_from_string": ([c_char_p], H),
    "kcnn_component_read": ([c_char_p, c_size_t, I], H),
    "kcnn_component_write": ([H, I, PP, PS], c_int),
    "kcnn_component_copy": ([H], H),
    "kcnn_component_delete": ([H], None),
    "kcnn_component_type": ([H], c_char_p),
    "kcnn_component_info": ([H, c_char_p, c_size_t], c_int),
    "kcnn_component_input_dim": ([H], c_int),
    "kcnn_component_output_dim": ([H], c_int),
    "kcnn_component_backprop_needs_input": ([H], c_int),
    "kcnn_component_backprop_needs_output": ([H], c_int),
    "kcnn_component_propagate": ([H, I] + MAT + MAT, c_int),
    "kcnn_component_backprop": ([H, I, P, I, P, I, P, I, I, H, P, I], c_int),
    "kcnn_component_propagate_chunks": ([H, I, I, I, I, I] + MAT + MAT, c_int),
    "kcnn_component_backprop_chunks": ([H, I, I, I, I, I, P, I, I, P, I], c_int),
    "kcnn_component_params": ([H, I, PP, PI, PI, PI], c_int),
    "kcnn_component_set_learning_rate": ([H, F], c_int),
    "kcnn_component_learning_rate": ([H], c_float),
    "kcnn_component_set_weight_decay_momentum": ([H, F, F], c_int),
    "kcnn_component_get_weight_decay_momentum": ([H, ctypes.POINTER(c_float), ctypes.POINTER(c_float)], c_int),
    "kcnn_component_set_index_routing": ([H, I], c_int),
    "kcnn_component_set_deferred_update": ([H, I], c_int),
    "kcnn_component_gradient_floats": ([H], c_size_t),
    "kcnn_component_set_gradient_storage": ([H, P], c_int),
    "kcnn_component_gradient": ([H, I, PP, PI, PI, PI], c_int),
    "kcnn_component_apply_gradient": ([H, I], c_int),
    "kcnn_nnet_new_from_config": ([c_char_p, I], H),
    "kcnn_nnet_read": ([c_char_p, c_size_t, I], H),
    "kcnn_nnet_write": ([H, I, PP, PS], c_int),
    "kcnn_nnet_delete": ([H], None),
    "kcnn_nnet_num_components": ([H], c_int),
    "kcnn_nnet_component": ([H, I], H),
    "kcnn_nnet_input_dim": ([H], c_int),
    "kcnn_nnet_output_dim": ([H], c_int),
    "kcnn_nnet_forward": ([H, P, I, I], c_int),
    "kcnn_nnet_forward_range": ([H, P, I, I, I, I], c_int),
    "kcnn_nnet_objf_and_deriv": ([H, P], c_int),
    "kcnn_nnet_backward": ([H, I, I], c_int),
    "kcnn_nnet_activation": ([H, I, PP, PI, PI, PI], c_int),
    "kcnn_nnet_input_deriv": ([H, PP, PI, PI, PI], c_int),
    "kcnn_nnet_objf_and_reset": ([H], ctypes.c_double),
    "kcnn_nnet_set_deferred_update": ([H, I], c_int),
    "kcnn_nnet_gradient_floats": ([H], c_size_t),
    "kcnn_nnet_set_gradient_arena": ([H, P], c_int),
    "kcnn_nnet_gradient_bucket": ([H, I, PS, PS], c_int),
    "kcnn_nnet_apply_gradients": ([H, I], c_int),
    "kcnn_nnet_train_minibatch_host": ([H, P, P, I, ctypes.POINTER(ctypes.c_double)], c_int),
    "kcnn_nnet_train_step": ([H, P, I, I, P], c_int),
    "kcnn_nnet_train_minibatch_host_async": ([H, P, P, I], c_int),
    "kcnn_nnet_running_objf": ([H], ctypes.c_double),
    "kcnn_nnet_last_step_replayed": ([H], c_int),
    "kcnn_nnet_set_fusion": ([H, I], c_int),
    "kcnn_nnet_fused_active": ([H], c_int),
    "kcnn_nnet_frames_per_example": ([H], c_int),
    "kcnn_p2p_flag_floats": ([], c_size_t),
    "kcnn_p2p_allreduce_f32": ([P, P, I, I, c_size_t, c_size_t, c_size_t, I], c_int),
    "kcnn_p2p_allreduce_multicast_f32": ([P, P, ctypes.c_ulonglong, I, I, c_size_t, c_size_t, c_size_t, I], c_int),
    "kcnn_p2p_error": ([P, c_size_t], c_int),
}


def declare(L):
    for name, (args, res) in PROTOS.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            continue
        fn.argtypes = args
        fn.restype = res
