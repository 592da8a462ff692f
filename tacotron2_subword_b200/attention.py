"""Attention parameter containers (state_dict-compatible with /root/reference/attention.py).

The arithmetic of both attentions runs inside the persistent CUDA decoder
(csrc/taco2dec.cu, phase B); these modules only own the weights under the reference's
names so ``load_state_dict(strict=True)`` works on reference checkpoints:

  StepwiseMonotonicAttention (attention.py:291-322): memory_layer.linear_layer.weight [A,E],
      v.weight [1,A] (a plain nn.Linear), query_layer.linear_layer.weight [A,H]
  LocationSensitiveAttention (attention.py:25-37): query_layer, memory_layer,
      v.linear_layer.weight [1,A] (a LinearNorm), location_layer.location_conv.conv.weight [F,2,K],
      location_layer.location_dense.linear_layer.weight [A,F]
"""
from torch import nn

from .layers import ConvNorm, LinearNorm

SMA = "StepwiseMonotonicAttention"
LSA = "LocationSensitiveAttention"


class LocationLayer(nn.Module):
    def __init__(self, attention_n_filters, attention_kernel_size, attention_dim):
        super().__init__()
        self.location_conv = ConvNorm(2, attention_n_filters, kernel_size=attention_kernel_size,
                                      padding=(attention_kernel_size - 1) // 2, bias=False)
        self.location_dense = LinearNorm(attention_n_filters, attention_dim, bias=False, w_init_gain="tanh")


class LocationSensitiveAttention(nn.Module):
    kind = LSA

    def __init__(self, attention_rnn_dim, embedding_dim, attention_dim, attention_location_n_filters,
                 attention_location_kernel_size):
        super().__init__()
        self.query_layer = LinearNorm(attention_rnn_dim, attention_dim, bias=False, w_init_gain="tanh")
        self.memory_layer = LinearNorm(embedding_dim, attention_dim, bias=False, w_init_gain="tanh")
        self.v = LinearNorm(attention_dim, 1, bias=False)
        self.location_layer = LocationLayer(attention_location_n_filters, attention_location_kernel_size,
                                            attention_dim)

    def v_weight(self):
        return self.v.linear_layer.weight


class StepwiseMonotonicAttention(nn.Module):
    kind = SMA
    sigmoid_noise = 2.0  # attention.py:316 (compiled into the kernel)

    def __init__(self, attention_rnn_dim, embedding_dim, attention_dim, attention_location_n_filters=None,
                 attention_location_kernel_size=None):
        super().__init__()
        self.memory_layer = LinearNorm(embedding_dim, attention_dim, bias=False, w_init_gain="tanh")
        self.v = nn.Linear(attention_dim, 1, bias=False)
        self.query_layer = LinearNorm(attention_rnn_dim, attention_dim, bias=False, w_init_gain="tanh")

    def v_weight(self):
        return self.v.weight
