;; bench/jvm/run_reference.clj — times the REAL reference path (Clojure + Fenzo on the JVM) on the
;; synthetic traces bench.py uses, for a box that has a JVM and Cook's ~/.m2 (this image has neither:
;; SURVEY §8c; until it runs, bench.py --impl reference times the C++ restatement and says so).
;;
;;   cd /path/to/Cook/scheduler && lein run -m clojure.main /path/to/repo/bench/jvm/run_reference.clj \
;;        /path/to/trace-dir c2
;;
;; trace-dir holds the columns `python -m cook_b200.traces --dump c2 DIR` writes (one little-endian
;; .bin per column + manifest.edn).  Modelled on test/cook/test/benchmark.clj:36-75 and
;; zz_simulator.clj:355-558: jobs are transacted into an in-memory Datomic, offers are Mesos-shaped maps.
(ns run-reference
  (:require [clojure.edn :as edn]
            [clojure.java.io :as io]
            [cook.scheduler.scheduler :as sched]
            [cook.rebalancer :as rebalancer]
            [cook.test.testutil :as testutil]
            [criterium.core :as crit]
            [datomic.api :as d])
  (:import (java.nio ByteBuffer ByteOrder)
           (java.nio.file Files)))

(defn- column [dir nm kind]
  (let [b (doto (ByteBuffer/wrap (Files/readAllBytes (.toPath (io/file dir (str nm ".bin"))))) (.order ByteOrder/LITTLE_ENDIAN))]
    (case kind
      :f64 (let [a (double-array (quot (.remaining b) 8))] (.get (.asDoubleBuffer b) a) a)
      :i32 (let [a (int-array (quot (.remaining b) 4))] (.get (.asIntBuffer b) a) a)
      :i64 (let [a (long-array (quot (.remaining b) 8))] (.get (.asLongBuffer b) a) a))))

(defn -main [dir config]
  (let [manifest (edn/read-string (slurp (io/file dir "manifest.edn")))
        conn (testutil/restore-fresh-database! (str "datomic:mem://bench-" (java.util.UUID/randomUUID)))
        user (column dir "pending_user" :i32) cpus (column dir "pending_cpus" :f64) mem (column dir "pending_mem" :f64)
        prio (column dir "pending_priority" :i32)
        _ (doseq [chunk (partition-all 1000 (range (alength user)))]
            (doseq [i chunk]
              (testutil/create-dummy-job conn :user (str "u" (aget user i)) :ncpus (aget cpus i) :memory (aget mem i)
                                         :priority (aget prio i))))
        offers (mapv (fn [i c m] {:id {:value (str "offer-" i)} :hostname (format "host-%06d" i) :slave-id {:value (str "s" i)}
                                  :resources [{:name "cpus" :type :value-scalar :scalar c}
                                              {:name "mem" :type :value-scalar :scalar m}]})
                     (range) (column dir "offer_cpus" :f64) (column dir "offer_mem" :f64))
        db (d/db conn)
        pool-name "no-pool"
        n (count offers)]
    (println "rank-jobs" (alength user) "pending:")
    (crit/quick-bench (doall (get (sched/rank-jobs db identity) pool-name)))
    (let [pending (get (sched/rank-jobs db identity) pool-name)
          fenzo (sched/make-fenzo-state 100000 nil 1.0)      ; good-enough-fitness 1.0: every VM evaluated (deterministic)
          considerable (take (:num-considerable manifest (alength user)) pending)]
      (println "match-offer-to-schedule" (count considerable) "x" n ":")
      (crit/quick-bench (sched/match-offer-to-schedule db fenzo considerable offers (atom {}) pool-name))
      (when (#{"c4" "c5"} config)
        (println "rebalance:")
        (crit/quick-bench (rebalancer/rebalance db identity pending {} {:max-preemption 128 :min-dru-diff 0.5 :safe-dru-threshold 1.0}))))
    (println "evals per match pass =" (* (alength user) n))
    (shutdown-agents)))

(apply -main *command-line-args*)
