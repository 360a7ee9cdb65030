"""Break the e2e step (host batch -> retrieved candidates) into phases."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_rag_b200 as G
from gnn_rag_b200 import batching, evaluate, ops, synthetic as S
dev = torch.device("cuda")
c = S.CONFIGS["cfg2"]
args = S.model_args("ReaRev", entity_dim=c["D"], num_iter=c["T"], num_ins=c["I"], num_gnn=c["K"], use_cuda=True)
torch.manual_seed(0)
m = G.ReaRev(dict(args), S.WEBQSP_NUM_ENTITY, S.WEBQSP_NUM_RELATION, S.WEBQSP_NUM_WORD).eval()
hb = S.make_batch(1, B=c["B"], N=c["N"], E=c["E"], with_weights=False)
pinned = batching.pin_batch(hb)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(8):
    t0 = T()
    db = batching.stage_batch(pinned, dev, S.WEBQSP_NUM_RELATION + 1)
    t1 = T()
    loss, pred, dist, _ = m(db)
    t2 = T()
    r, nb = evaluate.retrieve(dist, db, S.WEBQSP_NUM_ENTITY, args["eps"])
    t3 = T()
    print("iter %d: stage %.2f ms  forward %.2f ms  retrieve %.2f ms  total %.2f ms  (cands %d)" % (
        it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, sum(len(x) for x in r)))
# numpy batch (pageable) instead of pinned
for it in range(3):
    t0 = T(); db = batching.stage_batch(hb, dev, S.WEBQSP_NUM_RELATION + 1); t1 = T()
    print("pageable numpy stage: %.2f ms" % ((t1 - t0) * 1e3))
