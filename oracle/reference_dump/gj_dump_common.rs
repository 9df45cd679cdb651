// gj_dump_common.rs -- shared body of the reference-side dumpers (included with include!() by
// gj_dump_{nqueens,tsp,vrp}.rs).  NOT part of the product: it is compiled INSIDE a checkout of
// CameleoGrey/greyjack-solver-rust (see README.md in this directory) and runs the REFERENCE's own
// OOPScoreRequester / score calculators / Mover on inputs written by
// tests/golden/make_reference_inputs.py, so that the oracle (oracle/gj_oracle.c) and the CUDA path
// can be pinned against outputs of the reference itself (tests/test_reference_dump.py).
//
// The including file defines:
//   type ScoreT;                               the example's score struct
//   fn score_to_vec(s: &ScoreT) -> Vec<f64>;   its levels, hard first
//   fn build_requester(instance: &Value, incremental: bool) -> OOPScoreRequester<...>;

use std::collections::{HashMap, HashSet, VecDeque};
use std::fs;
use serde_json::{json, Value};
use greyjack::agents::metaheuristic_bases::mover::Mover;

fn f64_vec(v: &Value) -> Vec<f64> {
    v.as_array().unwrap().iter().map(|x| x.as_f64().unwrap()).collect()
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() != 3 {
        eprintln!("usage: {} <inputs.json> <outputs.json>", args[0]);
        std::process::exit(2);
    }
    let input: Value = serde_json::from_str(&fs::read_to_string(&args[1]).unwrap()).unwrap();
    let instance = &input["instance"];

    // ---- request_score_plain (oop_score_requester.rs:336-355) through the example's PSC ----------------
    let samples: Vec<Vec<f64>> = input["samples"].as_array().unwrap().iter().map(f64_vec).collect();
    let mut plain_requester = build_requester(instance, false);
    let variable_names = plain_requester.variables_manager.get_variables_names_vec();
    let plain: Vec<Vec<f64>> = plain_requester.request_score_plain(&samples).iter().map(score_to_vec).collect();

    // ---- request_score_incremental (:443-463) through the example's ISC ----------------------------------
    let base = f64_vec(&input["base"]);
    let deltas: Vec<Vec<(usize, f64)>> = input["deltas"].as_array().unwrap().iter().map(|d| {
        d.as_array().unwrap().iter().map(|p| {
            let p = p.as_array().unwrap();
            (p[0].as_u64().unwrap() as usize, p[1].as_f64().unwrap())
        }).collect()
    }).collect();
    let mut incr_requester = build_requester(instance, true);
    let incremental: Vec<Vec<f64>> = incr_requester.request_score_incremental(&base, &deltas).iter().map(score_to_vec).collect();

    // ---- Mover::do_move (mover.rs:98-128), incremental form, one move kind at a time ----------------------
    // The mover draws from entropy (StdRng::from_entropy per draw), so the moves cannot be replayed by
    // seed; what is dumped is (candidate, kind, changed columns, new values) -- the loader recovers the
    // chosen ids from the columns and checks the oracle's mover emits the same list for them.
    let mut moves: Vec<Value> = Vec::new();
    if let Some(m) = input.get("mover") {
        let candidate = f64_vec(&m["candidate"]);
        let n_moves = m["n_moves"].as_u64().unwrap() as usize;
        let rate = m["tabu_entity_rate"].as_f64().unwrap();
        let vm = &incr_requester.variables_manager;
        for kind in 0..6usize {
            let mut probas = vec![0.0; 6];
            probas[kind] = 1.0;
            let mut size_map: HashMap<String, usize> = HashMap::new();
            let mut sets_map: HashMap<String, HashSet<usize>> = HashMap::new();
            let mut deque_map: HashMap<String, VecDeque<usize>> = HashMap::new();
            let mut rates_map: HashMap<String, f64> = HashMap::new();
            for (name, ids) in vm.semantic_groups_map.iter() {
                // tabu_search_base.rs:115-121, TabuSearchBase::new (mutation_rate_multiplier None -> 0.0)
                size_map.insert(name.clone(), std::cmp::max((rate * (ids.len() as f64)).ceil() as usize, 1));
                sets_map.insert(name.clone(), HashSet::new());
                deque_map.insert(name.clone(), VecDeque::new());
                rates_map.insert(name.clone(), 0.0);
            }
            let mut mover = Mover::new(rate, size_map, sets_map, deque_map, rates_map, Some(probas));
            for _ in 0..n_moves {
                let (_, cols, vals) = mover.do_move(&candidate, vm, true);
                if let (Some(cols), Some(mut vals)) = (cols, vals) {
                    vm.fix_deltas(&mut vals, Some(cols.clone()));       // tabu_search_base.rs:124-132
                    moves.push(json!({"kind": kind, "columns": cols, "values": vals}));
                } else {
                    moves.push(json!({"kind": kind, "columns": null, "values": null}));
                }
            }
        }
    }

    let out = json!({"variable_names": variable_names, "plain": plain, "incremental": incremental, "moves": moves});
    fs::write(&args[2], serde_json::to_string(&out).unwrap()).unwrap();
    println!("wrote {}", args[2]);
}
