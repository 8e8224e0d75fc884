"""CPU: the cached task loader (clsurvey_b200/data.py) yields exactly the batches a reference-style
`DataLoader(dataset, batch_size, shuffle=...)` yields from the same host-generator state (src/methods/method.py:1058-1061),
including the ragged last batch and the generator state left behind."""
import torch

from clsurvey_b200 import data as cdata


def _ds(n=37):
    g = torch.Generator().manual_seed(1)
    return torch.utils.data.TensorDataset(torch.randn(n, 3, 4, 4, generator=g), torch.randint(0, 5, (n,), generator=g))


def test_cached_loader_matches_dataloader_order():
    ds = _ds()
    for shuffle in (False, True):
        for bs in (8, 37, 50):
            torch.manual_seed(123)
            ref = [(x.clone(), y.clone()) for x, y in torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=shuffle, num_workers=0)]
            ref2 = [(x.clone(), y.clone()) for x, y in torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=shuffle, num_workers=0)]
            state_ref = torch.get_rng_state()
            torch.manual_seed(123)
            ld = cdata.CachedLoader(ds, bs, shuffle, device="cpu")
            got = [(x.clone(), y.clone()) for x, y in ld]
            got2 = [(x.clone(), y.clone()) for x, y in ld]             # second epoch: next permutation of the same stream
            assert len(ld) == len(ref) == len(got)
            for (a, b), (c, d) in zip(ref + ref2, got + got2):
                assert torch.equal(a, c) and torch.equal(b, d)
            assert torch.equal(torch.get_rng_state(), state_ref)       # the host generator is left in the same state


def test_generic_dataset_is_materialised_once():
    class DS(torch.utils.data.Dataset):
        calls = 0

        def __len__(self):
            return 10

        def __getitem__(self, i):
            DS.calls += 1
            return torch.full((3, 2, 2), float(i)), i % 3, "path%d" % i     # (image, label, path) like ImagePathlist

    ds = DS()
    a = cdata.CachedLoader(ds, 4, False, device="cpu")
    n_calls = DS.calls
    b = cdata.CachedLoader(ds, 4, False, device="cpu")
    assert DS.calls == n_calls == 10
    xs = torch.cat([x for x, _ in b])
    assert torch.equal(xs[:, 0, 0, 0], torch.arange(10.0))
    assert [int(v) for _, y in a for v in y] == [i % 3 for i in range(10)]
