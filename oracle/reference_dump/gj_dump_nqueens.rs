// examples/nqueens/src/bin/gj_dump.rs -- see oracle/reference_dump/README.md of greyjack-b200.
#[path = "../domain/mod.rs"] mod domain;
#[path = "../cotwin/mod.rs"] mod cotwin;
#[path = "../score/mod.rs"] mod score;
#[path = "../persistence/mod.rs"] mod persistence;

use greyjack::cotwin::CotwinBuilderTrait;
use greyjack::domain::DomainBuilderTrait;
use greyjack::score_calculation::score_requesters::OOPScoreRequester;
use greyjack::score_calculation::scores::SimpleScore;
use persistence::cotwin_builder::{EntityVariants, UtilityObjectVariants};
use persistence::{CotwinBuilder, DomainBuilder};

type ScoreT = SimpleScore;
fn score_to_vec(s: &ScoreT) -> Vec<f64> { vec![s.simple_value] }

fn build_requester<'a>(instance: &serde_json::Value, incremental: bool)
    -> OOPScoreRequester<EntityVariants<'a>, UtilityObjectVariants, ScoreT> {
    // the scorers only see the row_id VARIABLES (the samples) and column_id = queen index; the seeded
    // start positions of the domain do not enter any score
    let n = instance["n_queens"].as_u64().unwrap();
    let domain = DomainBuilder::new(n, instance["seed"].as_u64().unwrap()).build_domain_from_scratch();
    let cotwin = CotwinBuilder::new(incremental).build_cotwin(domain, false);
    OOPScoreRequester::new(cotwin)
}

include!("gj_dump_common.rs");
