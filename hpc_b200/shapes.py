"""The named graph shapes (pure data, no imports: bench.py's CPU reference arm loads this file on its own, without the
package and therefore without the CUDA library).

Stand-ins for the OGB / DGL graphs the reference reads from ~/PA4/data (PA4/handout/script/run_all.sh:3-11): rows and
nnz from the public dataset cards, max row nnz from PA4/workspace/phase_2.log.
"""

# name -> (num_v, nnz, max_deg, tail_k, zero_ppm, local_ppm, window)
# max_deg: PA4/workspace/phase_2.log (the only per-graph statistic the reference pins). num_v / nnz: the public
# OGB / DGL / CogDL dataset cards as distributed with the course's data directory — they appear nowhere in the
# reference repository (discrepancy 4). tail_k / zero_ppm / local_ppm / window are this repository's
# choices for the synthetic stand-ins.
GRAPH_SHAPES = {
    # BASELINE.json configs
    "c0": (4096, 65536, 1024, 3, 50000, 300000, 64),
    "arxiv": (169343, 1166243, 13155, 3, 350000, 300000, 2048),          # phase_2.log:8
    "reddit": (232965, 114615892, 21657, 2, 0, 500000, 4096),            # phase_2.log:134 (reddit.dgl)
    "products": (2449029, 123718280, 17481, 2, 20000, 500000, 8192),     # phase_2.log:155
    # the other ten graphs of PA4/handout/script/run_all.sh:3
    "collab": (235868, 2358104, 671, 2, 50000, 400000, 2048),            # phase_2.log:29
    "citation": (2927963, 30387995, 1738, 2, 100000, 400000, 8192),      # phase_2.log:50
    "ddi": (4267, 2135822, 2234, 1, 0, 300000, 512),                     # phase_2.log:71
    "protein": (132534, 79122504, 7750, 2, 0, 500000, 4096),             # phase_2.log:92
    "ppa": (576289, 42463862, 3241, 2, 0, 500000, 4096),                 # phase_2.log:113
    "youtube": (1138499, 5980886, 28754, 3, 0, 300000, 4096),            # phase_2.log:176
    "amazon_cogdl": (1569960, 264339468, 75134, 2, 0, 500000, 8192),     # phase_2.log:197
    "yelp": (716847, 13954819, 4886, 2, 0, 400000, 4096),                # phase_2.log:218
    "wikikg2": (2500604, 16109182, 911, 2, 300000, 300000, 8192),        # phase_2.log:239
    "am": (881680, 5668682, 154828, 4, 300000, 300000, 4096),            # phase_2.log:260
}
# run_all.sh's dataset list, in its order ("reddit.dgl" is "reddit" here)
RUN_ALL_DATASETS = ("arxiv", "collab", "citation", "ddi", "protein", "ppa", "reddit", "products", "youtube",
                    "amazon_cogdl", "yelp", "wikikg2", "am")
