"""Host-side helpers mirroring ``endiffusion/models/utils.py`` plus mask canonicalisation.

The kernels take masks as ``sizes[b]`` (number of real nodes of molecule ``b``): the reference sampler
only ever builds prefix node masks and ``1 - eye`` edge masks (``diffusion_qm9.py:350-359``).  The
functions here convert the reference's dense bool masks to ``sizes`` and REFUSE anything else.
"""
import torch


def masks_from_sizes(sizes, N, device):
    """diffusion_qm9.py:350-359: node_mask [B,N,1] bool, edge_mask [B,N,N] bool from per-molecule sizes."""
    s = torch.as_tensor(sizes, device=device).view(-1, 1)
    ar = torch.arange(N, device=device).view(1, -1)
    nm = ar < s
    em = nm.unsqueeze(2) & nm.unsqueeze(1) & ~torch.eye(N, dtype=torch.bool, device=device).unsqueeze(0)
    return nm.unsqueeze(2), em


def sizes_from_node_mask(node_mask, B, N):
    """[B,N,1] / [B*N,1] mask -> int32 sizes [B]; raises unless every molecule's real nodes are a prefix."""
    nm = node_mask.reshape(B, N) != 0
    sizes = nm.sum(1).to(torch.int32)
    expect = torch.arange(N, device=nm.device).view(1, -1) < sizes.view(-1, 1)
    if not bool((nm == expect).all()):
        raise NotImplementedError("node_mask must mark a prefix of each molecule's nodes "
                                  "(the layout the reference sampler builds); other masks are not built")
    return sizes.contiguous()


def check_edge_mask(edge_mask, sizes, B, N):
    _, em = masks_from_sizes(sizes, N, edge_mask.device)
    if not bool(((edge_mask.reshape(B, N, N) != 0) == em).all()):
        raise NotImplementedError("edge_mask must be (node_mask x node_mask) minus the diagonal "
                                  "(diffusion_qm9.py:357); other edge masks are not built")


def check_edge_index(edge_index, B, N):
    """The canonical dense list of en_dynamics.py:131-136 (b-major, i-major, j-minor)."""
    rows, cols = edge_index
    E = B * N * N
    if rows.numel() != E or cols.numel() != E:
        raise NotImplementedError("edge_index must be the dense all-pairs list of get_adj_matrix")
    e = torch.arange(E, device=rows.device)
    b = e // (N * N)
    if not (bool((rows == b * N + (e // N) % N).all()) and bool((cols == b * N + e % N).all())):
        raise NotImplementedError("edge_index must be the dense all-pairs list of get_adj_matrix")


def sizes_from_masks(node_mask, edge_mask, edge_index, BN):
    """Recover (B, N, sizes) from the reference's EGNN.forward arguments, validating them on every call (one small
    reduction; the sampling loop passes ``sizes`` directly and never comes here)."""
    if node_mask is None or edge_mask is None:
        raise NotImplementedError("node_mask and edge_mask are required (the sampler always passes them)")
    E = edge_mask.numel()
    if E % BN:
        raise ValueError("edge_mask does not match h")
    N = E // BN
    B = BN // N
    sizes = sizes_from_node_mask(node_mask, B, N)
    check_edge_mask(edge_mask, sizes, B, N)
    if edge_index is not None:
        check_edge_index(edge_index, B, N)
    return B, N, sizes


def remove_mean_with_mask(x, node_mask):
    """models/utils.py:43-57 (host/torch version, used outside the captured loop)."""
    masked = (x * (~node_mask)).abs().sum().item()
    assert masked < 1e-5, f"Error {masked} too high"
    n = node_mask.sum(1, keepdim=True)
    return x - (x.sum(1, keepdim=True) / n) * node_mask


def assert_correctly_masked(variable, node_mask):
    """models/utils.py:72-75."""
    assert (variable * ~node_mask).abs().max().item() < 1e-4, "Variables not masked properly."


def assert_mean_zero_with_mask(x, node_mask, eps=1e-10):
    """models/utils.py:65-70."""
    assert_correctly_masked(x, node_mask)
    largest = x.abs().max().item()
    err = x.sum(1, keepdim=True).abs().max().item()
    rel = err / (largest + eps)
    assert rel < 1e-2, f"Mean is not zero, relative_error {rel}"
