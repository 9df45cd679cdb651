// examples/vrp/src/bin/gj_dump.rs
// -- see oracle/reference_dump/README.md of greyjack-b200.
#[path = "../domain/mod.rs"] mod domain;
#[path = "../cotwin/mod.rs"] mod cotwin;
#[path = "../score/mod.rs"] mod score;
#[path = "../persistence/mod.rs"] mod persistence;

use greyjack::cotwin::CotwinBuilderTrait;
use greyjack::domain::DomainBuilderTrait;
use greyjack::score_calculation::score_requesters::OOPScoreRequester;
use greyjack::score_calculation::scores::HardMediumSoftScore;
use persistence::cotwin_builder::{EntityVariants, UtilityObjectVariants};
use persistence::{CotwinBuilder, DomainBuilder};

type ScoreT = HardMediumSoftScore;
fn score_to_vec(s: &ScoreT) -> Vec<f64> { vec![s.hard_score, s.medium_score, s.soft_score] }

fn build_requester<'a>(instance: &serde_json::Value, incremental: bool)
    -> OOPScoreRequester<EntityVariants<'a>, UtilityObjectVariants, ScoreT> {
    let domain = DomainBuilder::new(instance["path"].as_str().unwrap()).build_domain_from_scratch();
    let cotwin = CotwinBuilder::new(incremental, false).build_cotwin(domain, false);
    OOPScoreRequester::new(cotwin)
}

include!("gj_dump_common.rs");
