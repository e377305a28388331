"""B200-native (sm_100a) TTT hot path: drop-in for the reference's ttt-tk op.

Public surface mirrors the reference:
  test_time_training.ttt_forward / ttt_backward      <-> ttt-tk/test_time_training.cpp:95-105
  mlp_tk.TkMLP (torch.autograd.Function)             <-> ttt/models/ssm/mlp_tk.py:9
"""
__all__ = ["_lib", "test_time_training", "mlp_tk", "linear_triton", "seq_block", "seq_shard", "attention", "process_input",
           "ttt_layer", "interleave", "host_stream", "rope", "transformer_layer"]
