"""End-to-end quality run of the UNMODIFIED reference on the CPU (build container only): the loop of the reference's
main.py (143-194) for --model svd, using the reference's own train_SDRM / sample_ddpm / compute_mf_results.
Writes one JSON line per run with Recall@10 of F-SDRM / M-SDRM / MultiVAE++."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from oracle import refstub
ref = refstub.import_reference()
import dataloaders as rd, svd_benchmark as rsvd  # reference modules
from torch.utils.data import DataLoader

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 5
only_synth = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
out_path = sys.argv[3] if len(sys.argv) > 3 else "profiles/e2e_reference_ml100k.jsonl"
A = dict(T=83, B=550, H=930, L=830, nh=2, nd=1.0, epochs=265, lr=2.1e-5, vae_bs=780, vae_lr=6e-4)  # README trial 223
if len(sys.argv) > 4:
    A.update(json.loads(sys.argv[4]))
TRAIN, TPV, VALID = rd.load_data("ml-100k", "/root/reference/data")
N_USERS, N_ITEMS = TRAIN.shape
SPARSITY = 1 - (TRAIN.nnz / (N_USERS * N_ITEMS))
ds = rd.SparseDataset(TPV, TPV)
sampler = torch.utils.data.sampler.BatchSampler(torch.utils.data.sampler.RandomSampler(ds, generator=torch.Generator(device="cpu")), batch_size=A["B"], drop_last=False)
dl = DataLoader(ds, batch_size=1, collate_fn=rd.sparse_batch_collate, generator=torch.Generator(device="cpu"), sampler=sampler, shuffle=False)
for run in range(runs):
    t0 = time.time()
    DIFF, VAE = ref.train_SDRM(dl=dl, N_ITEMS=N_ITEMS, VAE_LATENT=A["L"], VAE_HIDDEN=A["H"], VAE_LR=A["vae_lr"], VAE_BATCH_SIZE=A["vae_bs"],
                               DIFF_LATENT=A["L"], DIFF_TRAINING_EPOCHS=A["epochs"], DIFF_LR=A["lr"], N_HIDDEN_MLP_LAYERS=A["nh"], TIMESTEPS=A["T"],
                               noise_divider=A["nd"], VAE_DIR_PATH="/tmp/ref_temp_vae", TRAIN_PARTIAL_VALID_DATA=TPV, VALID_DATA=VALID,
                               OPTIMIZATION_OBJECTIVE="Recall@10", verbose=False)
    t_train = time.time() - t0
    res = {"impl": "reference-cpu", "run": run, "train_s": round(t_train, 1), "only_synthetic": only_synth}
    t1 = time.time()
    M = ref.sample_ddpm(N_USERS, DIFF, VAE, A["L"], A["nd"], timesteps="random", n_timesteps=A["T"]).detach().cpu().numpy()
    res["sample_random_s"] = round(time.time() - t1, 2); t1 = time.time()
    F = ref.sample_ddpm(N_USERS, DIFF, VAE, A["L"], A["nd"], n_timesteps=A["T"]).detach().cpu().numpy()
    res["sample_full_s"] = round(time.time() - t1, 2)
    V = VAE.sample(N_USERS)
    for name, S in (("F-SDRM", F), ("M-SDRM", M), ("MultiVAE++", V)):
        syn = pd.DataFrame((S >= np.quantile(S.flatten(), SPARSITY)).astype(int))
        rec, ndcg = rsvd.compute_mf_results(TRAIN, VALID, synthetic_data=syn, nnmf=False, only_synthetic=only_synth)
        res[name] = {"recall@10": float(rec[3]), "ndcg@10": float(ndcg[3])}
    print(json.dumps(res), flush=True)
    with open(out_path, "a") as fh:
        fh.write(json.dumps(res) + "\n")
