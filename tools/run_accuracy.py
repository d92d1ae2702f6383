"""Accuracy parity runs against BASELINE.md §2.2 (reference: cfg1 Sup MUTAG 69.0 +- 11.6 % final-epoch test accuracy
over folds 0-9; cfg2 Unsup PTC degree-as-tag 87.8 +- 7.0 % after 50 epochs)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import numpy as np
import train_pytorch_U2GNN_Sup as SUP
import train_pytorch_U2GNN_UnSup as UNSUP

which = sys.argv[1] if len(sys.argv) > 1 else "sup"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 50
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"
quiet = lambda *a, **k: None
if which == "sup":
    finals, best = [], []
    t0 = time.time()
    for fold in range(10):
        args = SUP.parse_args(["--dataset", "MUTAG", "--fold_idx", str(fold), "--num_neighbors", "8", "--num_timesteps", "3",
                               "--ff_hidden_size", "1024", "--batch_size", "4", "--num_epochs", str(epochs), "--precision", precision,
                               "--run_folder", "/tmp/u2gnn_runs/x/", "--model_name", "MUTAG_f%d" % fold])
        accs, _ = SUP.run(args, log=quiet)
        finals.append(accs[-1] * 100); best.append(max(accs) * 100)
        print("fold %d final %.2f best %.2f" % (fold, finals[-1], best[-1]), flush=True)
    print(json.dumps({"run": "cfg1 Sup MUTAG attn_axis=nodes " + precision, "epochs": epochs, "final_acc_per_fold": finals,
                      "mean": float(np.mean(finals)), "stdev": float(np.std(finals, ddof=1)), "reference": "69.0 +- 11.6",
                      "seconds": time.time() - t0}))
else:
    t0 = time.time()
    args = UNSUP.parse_args(["--dataset", "PTC", "--degree_as_tag", "--num_neighbors", "4", "--num_timesteps", "2", "--ff_hidden_size", "1024",
                             "--sampled_num", "512", "--batch_size", "2", "--learning_rate", "0.0001", "--num_epochs", str(epochs), "--precision", precision,
                             "--run_folder", "/tmp/u2gnn_runs/x/", "--model_name", "PTC_unsup"])
    hist = UNSUP.run(args, log=quiet)
    print(json.dumps({"run": "cfg2 Unsup PTC degree-as-tag attn_axis=nodes " + precision, "epochs": epochs, "epoch1": hist[0][:2],
                      "final": hist[-1][:2], "best": max(h[0] for h in hist), "reference": "epoch1 65.1 +- 6.1, epoch50 87.8 +- 7.0",
                      "seconds": time.time() - t0}))
