"""Host restatement of csrc/dp.cu (pack / unpack / loss from row statistics) used by the CPU gloo tests."""
import torch


class TorchExchangeMixin:
    """pack / unpack / loss_from_stats in plain torch for the gloo tests' injected backend (test infrastructure; the product
    backend is csrc/dp.cu): same bit-level packing."""

    @staticmethod
    def pack(emb, labels):
        lab = labels.contiguous().view(-1).to(torch.int64).view(torch.int32).view(-1, 2)        # little endian: low word first
        return torch.cat([emb, lab.view(torch.float32).to(emb.dtype)], dim=1) if emb.dtype == torch.float32 else \
            torch.cat([emb, labels.view(-1, 1).to(emb.dtype), torch.zeros_like(labels).view(-1, 1).to(emb.dtype)], dim=1)

    @staticmethod
    def unpack(packed, D):
        F = packed[:, :D].contiguous()
        if packed.dtype == torch.float32:
            y = packed[:, D:D + 2].contiguous().view(torch.int32).view(-1).view(torch.int64)
        else:                                   # fp64 test tensors carry the label value itself
            y = packed[:, D].to(torch.int64)
        return F, y

    @staticmethod
    def loss_from_stats(stats_all, temperature, base_temperature):
        npos, spos, den = stats_all[:, 2], stats_all[:, 3], stats_all[:, 1]
        nn_ = torch.where(npos == 0, torch.ones_like(npos), npos)
        row = -(temperature / base_temperature) * (spos - npos * torch.log(den)) / nn_
        return (row.double().sum() / stats_all.shape[0]).to(stats_all.dtype).reshape(1)
