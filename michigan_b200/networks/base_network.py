"""BaseNetwork — same surface as the reference's models/networks/base_network.py:10-59
(`print_network`, `init_weights(init_type, gain)`, static `modify_commandline_options`)."""
import torch.nn as nn
from torch.nn import init


class BaseNetwork(nn.Module):
    def __init__(self):
        super().__init__()

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def print_network(self):
        n = sum(p.numel() for p in self.parameters())
        print("Network [%s] was created. Total number of parameters: %.1f million. "
              "To see the architecture, do print(network)." % (type(self).__name__, n / 1000000))

    def init_weights(self, init_type="normal", gain=0.02):
        """Same visiting order and per-class rules as base_network.py:28-59, so a given torch seed yields
        the same initial weights as the reference (spectral-normed convs are reached through the
        `.weight` alias of `weight_orig`, see torch.nn.utils.spectral_norm)."""
        def init_func(m):
            name = m.__class__.__name__
            if name.find("BatchNorm2d") != -1:
                if getattr(m, "weight", None) is not None:
                    init.normal_(m.weight.data, 1.0, gain)
                if getattr(m, "bias", None) is not None:
                    init.constant_(m.bias.data, 0.0)
            elif hasattr(m, "weight") and (name.find("Conv") != -1 or name.find("Linear") != -1):
                if init_type == "normal":
                    init.normal_(m.weight.data, 0.0, gain)
                elif init_type == "xavier":
                    init.xavier_normal_(m.weight.data, gain=gain)
                elif init_type == "xavier_uniform":
                    init.xavier_uniform_(m.weight.data, gain=1.0)
                elif init_type == "kaiming":
                    init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
                elif init_type == "orthogonal":
                    init.orthogonal_(m.weight.data, gain=gain)
                elif init_type == "none":
                    m.reset_parameters()
                else:
                    raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
                if getattr(m, "bias", None) is not None:
                    init.constant_(m.bias.data, 0.0)

        self.apply(init_func)
        for m in self.children():
            if hasattr(m, "init_weights"):
                m.init_weights(init_type, gain)
