"""Host-side restatement of the conv kernel's stream-K work decomposition (csrc/conv_common.cuh: WorkList,
streamk_cta_of_unit, the owner's `parts` loop) checked for the properties the device protocol relies on:
every (tile, chunk) unit is reduced exactly once; a CTA is a helper at most once (its FIRST item) and an owner at most
once (its LAST item); the owner's expected helper count equals the CTAs that hold the rest of its tile, exactly the CTAs
that compute `owner_cta` == it; helpers are the CTAs immediately after the owner (their partial slots are indexed by CTA)."""
import pytest


def unit_begin(c, per, extra):
    return c * per + min(c, extra)


def cta_of_unit(u, per, extra):
    cut = extra * (per + 1)
    return u // (per + 1) if u < cut else extra + (u - cut) // per


def items_of(c, g, tiles, k):
    units = tiles * k
    per, extra = divmod(units, g)
    cur, end = unit_begin(c, per, extra), unit_begin(c + 1, per, extra)
    out = []
    while cur < end:
        tile = cur // k
        kb = cur - tile * k
        ke = min(k, kb + end - cur)
        out.append((tile, kb, ke))
        cur += ke - kb
    return out


@pytest.mark.parametrize("tiles,k,sms", [(224, 8, 148), (56, 8, 148), (112, 4, 148), (16, 8, 148), (810, 2, 148),
                                         (3, 8, 148), (1, 2, 148), (149, 3, 148), (224, 8, 132), (37, 5, 64)])
def test_streamk_schedule_properties(tiles, k, sms):
    units = tiles * k
    g = min(sms, units)
    per, extra = divmod(units, g)
    assert per >= 1
    covered = {}
    owners, helpers = {}, {}
    for c in range(g):
        its = items_of(c, g, tiles, k)
        assert its, "every CTA of the grid has work"
        for j, (tile, kb, ke) in enumerate(its):
            assert 0 <= kb < ke <= k
            for u in range(tile * k + kb, tile * k + ke):
                assert u not in covered
                covered[u] = c
            if kb > 0:                                    # helper part
                assert j == 0, "a helper part is the FIRST item of its CTA"
                assert c not in helpers
                helpers[c] = (tile, cta_of_unit(tile * k, per, extra))
            elif ke < k:                                  # owner
                assert j == len(its) - 1, "the owned partial tile is the LAST item of its CTA"
                tile_end = (tile + 1) * k
                parts, cc = 0, c + 1
                while unit_begin(cc, per, extra) < tile_end:
                    cc += 1
                    parts += 1
                owners[c] = (tile, parts)
    assert len(covered) == units
    for c, (tile, parts) in owners.items():
        mine = sorted(h for h, (t, o) in helpers.items() if t == tile)
        assert mine == list(range(c + 1, c + 1 + parts)), (c, tile, parts, mine)
        assert all(helpers[h][1] == c for h in mine)
    for h, (tile, o) in helpers.items():
        assert o in owners and owners[o][0] == tile
    # balance: no CTA reduces more than one chunk above the mean
    loads = [sum(ke - kb for _, kb, ke in items_of(c, g, tiles, k)) for c in range(g)]
    assert max(loads) - min(loads) <= 1
