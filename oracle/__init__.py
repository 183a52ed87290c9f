"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement + reference build). Never imported by spring_b200."""
