// examples/vrp_service/src/bin/gj_dump.rs -- see oracle/reference_dump/README.md of greyjack-b200.
// (vrp_service builds its domain from a JSON Value instead of a .vrp file,
// persistence/domain_builder.rs:20-110.)
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
    let text = std::fs::read_to_string(instance["path"].as_str().unwrap()).unwrap();
    let vrp_json: serde_json::Value = serde_json::from_str(&text).unwrap();
    let domain = DomainBuilder::new(&vrp_json).build_domain_from_scratch();
    let cotwin = CotwinBuilder::new(incremental, false).build_cotwin(domain, false);
    OOPScoreRequester::new(cotwin)
}

include!("gj_dump_common.rs");
