"""Host-side logic of the peer-memory data-parallel path that needs no GPU: region layout, the bucket cut points the backward
generator promises, and the C-ABI's host-only queries."""
import torch


def test_peer_region_layout_on_host():
    from phoneme_contrast_b200 import _lib as L
    from phoneme_contrast_b200.peer import PeerRegion
    N, D, P = 48, 128, 4 * 1001
    fields = [("F", (N, D), torch.float32), ("y", (N,), torch.int64), ("stats", (N, 4), torch.float32), ("flat", (P,), torch.float32)]
    probe = PeerRegion(fields, "cpu", bases=[16, 32], rank=1, buf=None)
    flag = L.lib().pc_peer_flag_bytes()
    assert flag % 256 == 0 and flag >= 4 * (2 * L.lib().pc_peer_max_ranks() + 3)
    offs = [probe.offset(n) for n, _, _ in fields]
    assert offs[0] >= flag and all(o % 256 == 0 for o in offs) and offs == sorted(offs)
    sizes = [N * D * 4, N * 8, N * 16, P * 4]
    for o, sz, nxt in zip(offs, sizes, offs[1:] + [probe.nbytes]):
        assert o + sz <= nxt                        # fields do not overlap
    assert probe.world_size == 2 and probe.rank == 1
    buf = torch.zeros(probe.nbytes, dtype=torch.uint8)
    reg = PeerRegion(fields, "cpu", bases=[16, 32], rank=0, buf=buf)
    f, y = reg.local("F"), reg.local("y")
    assert f.shape == (N, D) and f.dtype == torch.float32 and y.shape == (N,) and y.dtype == torch.int64
    f.fill_(1.0)
    assert int(buf[reg.offset("F"):reg.offset("F") + 4].view(torch.float32)[0]) == 1 and int(reg.local("y").abs().sum()) == 0


def test_tail_bucket_offsets_follow_parameter_order():
    from phoneme_contrast_b200.models import model_registry
    for arch, cfg in (("phoneme_cnn_deep", {"hidden_dims": [64, 64, 128, 128]}), ("phoneme_cnn", {})):
        m = model_registry.create(arch, cfg)
        offs = m.tail_bucket_offsets()
        assert len(offs) == 2 and offs[0] > offs[1] > 0 and offs[0] == m.tail_bucket_offset()
        params = list(m.parameters())
        starts, off = {}, 0
        for p in params:
            starts[id(p)] = off
            off += p.numel()
        nb = len(m.conv_blocks)
        for k, bi in enumerate((nb - 1, nb - 2)):
            assert offs[k] == starts[id(next(m.conv_blocks[bi].parameters()))]
        # everything behind offs[0] = last block + attention + projection: the part the backward completes first
        tail = sum(p.numel() for p in m.conv_blocks[nb - 1].parameters()) + sum(p.numel() for p in m.projection.parameters())
        tail += sum(p.numel() for p in m.attention.parameters()) if m.use_attention else 0
        assert off - offs[0] == tail
